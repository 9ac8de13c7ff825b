"""`python -m multithreading_string_matching_b200 <file.pcap> <string.txt> [n] [udp|tcp]` -> bin/kmp_match."""
import os
import sys

exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", "kmp_match")
if not os.path.isfile(exe):
    sys.exit("kmp_match is not built: make -C %s" % os.path.dirname(exe))
os.execv(exe, [exe] + sys.argv[1:])
