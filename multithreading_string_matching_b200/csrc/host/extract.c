/* extract.c -- which bytes of a frame are "the payload": the host-side input contract of the path.
 *
 * Byte-faithful to the reference's two extractors, including what they do NOT check:
 *   dump_UDP_packet  packet_dumping.h:87-139   no ethertype / IP version / IHL>=5 / UDP length test;
 *                                              Ethernet trailer padding is part of the payload
 *   dump_TCP_packet  packet_dumping.h:150-188  no protocol test at all
 * Frames too short for the headers they announce make the reference's TCP extractor read outside
 * the frame and wrap an unsigned length (:162,172,182); those are rejected here.
 */
#include "kmpb_internal.h"

#define ETHER_BYTES 14u   /* sizeof(struct ether_header), :94 */
#define IPV4_BYTES 20u    /* sizeof(struct ip), :102 */
#define UDP_BYTES 8u      /* sizeof(struct UDP_hdr) for the gate at :125 and sizeof(pointer) at :133 */
#define TCP_MIN_BYTES 20u
#define PROTO_FIELD 9u    /* offset of ip_p inside the IPv4 header */
#define TCP_DOFF_FIELD 12u

int kmpb_extract_udp(const uint8_t *frame, uint32_t frame_len, uint32_t *payload_off, uint32_t *payload_len)
{
    if (frame_len < ETHER_BYTES) return 0;
    uint32_t left = frame_len - ETHER_BYTES;
    if (left < IPV4_BYTES) return 0;
    const uint8_t *ip = frame + ETHER_BYTES;
    uint32_t ip_bytes = 4u * (ip[0] & 0x0fu);
    if (left < ip_bytes) return 0;
    if (ip[PROTO_FIELD] != 17) return 0;
    left -= ip_bytes;
    if (left < UDP_BYTES) return 0;
    *payload_off = ETHER_BYTES + ip_bytes + UDP_BYTES;
    *payload_len = left - UDP_BYTES;
    return 1;
}

int kmpb_extract_tcp(const uint8_t *frame, uint32_t frame_len, uint32_t *payload_off, uint32_t *payload_len)
{
    if (frame_len <= ETHER_BYTES) return 0;
    uint32_t ip_bytes = 4u * (frame[ETHER_BYTES] & 0x0fu);
    if (ip_bytes < IPV4_BYTES) return 0;
    uint32_t tcp_at = ETHER_BYTES + ip_bytes;
    if (frame_len <= tcp_at + TCP_DOFF_FIELD) return 0;
    uint32_t tcp_bytes = 4u * (uint32_t)(frame[tcp_at + TCP_DOFF_FIELD] >> 4);
    if (tcp_bytes < TCP_MIN_BYTES) return 0;
    if (frame_len < tcp_at + tcp_bytes) return 0;
    *payload_off = tcp_at + tcp_bytes;
    *payload_len = frame_len - (tcp_at + tcp_bytes);
    return 1;
}
