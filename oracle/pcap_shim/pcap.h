/*
 * pcap.h -- minimal stand-in for libpcap's savefile API.  TEST INFRASTRUCTURE ONLY.
 *
 * libpcap is not installed in this image (no pcap.h, no libpcap.so), and it is an un-vendored,
 * un-pinned dependency of the reference (the only mention is "-lpcap" in the header comments of
 * /root/reference/serial.c:2 and openmp_data.c:1).  libpcap does record framing only: none of the
 * hot path's arithmetic lives in it.  This shim implements exactly the four symbols the reference's
 * file-based programs call (pcap_open_offline serial.c:91, pcap_next_ex serial.c:115, pcap_close
 * mpi_dumping.c:130, plus struct pcap_pkthdr / PCAP_ERRBUF_SIZE / bpf_u_int32) so that the
 * UNMODIFIED reference sources compile into oracle/_ref/ (see oracle/Makefile).
 *
 * Savefile format implemented: classic pcap (magic a1b2c3d4 usec / a1b23c4d nsec, either byte
 * order), 24-byte global header, 16-byte record headers.  pcapng is not supported.
 */
#ifndef ORACLE_PCAP_SHIM_H
#define ORACLE_PCAP_SHIM_H

#include <stdio.h>
#include <sys/types.h>
#include <sys/time.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCAP_ERRBUF_SIZE 256
#define PCAP_ERROR (-1)
#define PCAP_ERROR_BREAK (-2)

typedef unsigned int bpf_u_int32;
typedef int bpf_int32;

struct pcap_pkthdr {
    struct timeval ts;  /* time stamp */
    bpf_u_int32 caplen; /* length of portion present in the file */
    bpf_u_int32 len;    /* length of this packet on the wire */
};

typedef struct pcap pcap_t;

pcap_t *pcap_open_offline(const char *fname, char *errbuf);
/* returns 1 per record, PCAP_ERROR_BREAK (-2) at end of file, PCAP_ERROR (-1) on a damaged file */
int pcap_next_ex(pcap_t *p, struct pcap_pkthdr **pkt_header, const u_char **pkt_data);
void pcap_close(pcap_t *p);
char *pcap_geterr(pcap_t *p);
int pcap_datalink(pcap_t *p);
int pcap_snapshot(pcap_t *p);

#ifdef __cplusplus
}
#endif
#endif
