// union_kernel.cu -- the union engine (KMPB_ENGINE_UNION): every payload byte is read from HBM once.
//
// Two levels, strictly separated.
//
//  FAST PATH (every byte) knows nothing about packets.  A warp streams work items -- runs of whole
//  packets, ~64 KB of the flat CSR byte buffer -- in rows of 1024 contiguous bytes.  Rows travel
//  global -> shared memory by per-lane 16-byte asynchronous copies (cp.async, SASS LDGSTS) into a
//  per-warp ring of UN_SLOTS slots, one commit group per row; a slot holds the row plus the 16 bytes
//  after it, so every lane finds its lookahead in its own slot.  Each lane pushes its 32 bytes (+3
//  bytes of lookahead) through a 4-byte-deep shift-and filter over 8 buckets:
//      S = ((S << 8) | 0xff) & filter[byte]
//  filter[] lives in shared memory in a bank-private layout (byte address = byte*256 + lane*4) at a
//  64 KB-aligned shared address, so the one lookup per byte never bank-conflicts and its complete
//  address is a single PRMT of the text word with a per-lane constant.  The shift-or-0xff is one
//  integer multiply-add (FMA pipe), the AND one LOP3 (ALU pipe).  Bits 24..30 of S say "the last 4
//  bytes are the first 4 bytes (or all the bytes) of some pattern of bucket b"; bit 31 says "this
//  byte is NUL".  The reports are OR-ed per quarter of the group (8 start positions); a lane with any
//  report appends an EVENT -- its 32 bytes and the 8 after them, where they are, and which quarters reported -- to
//  the warp's list in shared memory.  That is all: one ballot and (usually) a few stores per row on top of the
//  filter.  (When the list is full and a row reports from many groups -- NUL-dense payloads -- events that hold
//  only NULs and are superseded by a later one of the same row are dropped first: drop_superseded.)
//
//  SLOW PATH (events only).  When the list cannot take the next row's events the warp resolves up to 32
//  of them at once, in stream order.
//  Phase 1, one event per lane:
//    - the lane re-runs the filter over the quarters that reported, this time recording which start
//      positions fired and which bytes are NUL;
//    - it finds the packet that holds its first candidate in its item's slice of `offsets` (interpolation
//      guess, then binary search); the item's packet and byte range waits in the warp's scratch words;
//    - a candidate start q in packet [ps, pe) is alive when no NUL lies in [ps, q) -- the reference's
//      "text ends at the first NUL" rule (serial.c:191).  NULs inside the group come from the lane's own
//      mask; the last NUL before the group comes from the nearest earlier event that held one (events
//      are in stream order, every NUL byte of the stream raises one) or from the warp's carry.
//  Phase 2, one alive candidate per lane, whichever event it came from (they are numbered across the
//  lanes by a prefix sum): its first two bytes select a slot of the verification tables of automaton.c, the
//  slot's pattern records (first 8 bytes + masks, length, id) are compared, longer patterns word by word,
//  and a hit is counted when it ends inside its packet (q + len <= pe).
//  So every pattern occurrence that lies inside one packet and has no NUL before it in that packet is
//  counted exactly once.  Counts go to shared-memory counters and leave the block as one atomic per
//  distinct pattern.  No separators, no padding and no second pass over the payload.
//
//  Measured and dropped: a TMA bulk-copy ring (1 KB cp.async.bulk per row, 3 % slower:
//  ~25 instructions per row to issue one copy from one elected lane), direct 16-byte loads with an L2
//  prefetch (10 % slower: exposed latency), 64 bytes per lane (17 % slower), a chunk-major slot layout.
#include <algorithm>

#include "kmpb_device.cuh"

#ifndef KMPB_UN_THREADS
#define KMPB_UN_THREADS 896
#endif
#ifndef KMPB_UN_ITEM_KB
#define KMPB_UN_ITEM_KB 64
#endif
#ifndef KMPB_UN_SLOTS
#define KMPB_UN_SLOTS 2
#endif
constexpr int UN_THREADS = KMPB_UN_THREADS; // one block per SM
constexpr int UN_WARPS = UN_THREADS / 32;
constexpr uint32_t UN_GRP = 32;                           // bytes per lane per row
constexpr uint32_t UN_ROW = 32 * UN_GRP;                  // bytes per warp row
constexpr uint32_t UN_SLOTS = KMPB_UN_SLOTS;              // rows in flight per warp (cp.async -> shared memory)
// KMPB_UN_SLOTS == 2: a slot holds a row and the 16 bytes after it (lane 31's lookahead, copied by lane 31).
// KMPB_UN_SLOTS == 3: a slot holds a row; two rows are complete when a row is scanned, and lane 31 finds its
// lookahead at the start of the next slot (no tail copy).
#ifndef KMPB_UN_TAIL
#define KMPB_UN_TAIL (KMPB_UN_SLOTS < 3)
#endif
constexpr bool UN_TAIL = KMPB_UN_TAIL;
constexpr uint32_t UN_SLOT_BYTES = UN_ROW + (UN_TAIL ? 16 : 0);
constexpr uint32_t UN_ITEM_BYTES = KMPB_UN_ITEM_KB << 10; // target work-item size
#ifndef KMPB_UN_TAIL_ITEMS
#define KMPB_UN_TAIL_ITEMS 16384
#endif
constexpr uint32_t UN_TAIL_ITEMS = KMPB_UN_TAIL_ITEMS;   // small items at the end of a batch (about one per warp x 4)
#ifndef KMPB_UN_QCAP
#define KMPB_UN_QCAP 32
#endif
constexpr uint32_t UN_QCAP = KMPB_UN_QCAP;                          // events per warp list
constexpr uint32_t UN_Q_WORDS = 12; // event: 32 B group, 8 B lookahead, group index | item parity << 31, quarter reports (48 B)
#ifndef KMPB_UN_DENSE
#define KMPB_UN_DENSE 8
#endif
constexpr uint32_t UN_DENSE = KMPB_UN_DENSE; // reporting groups per row from which superseded NUL-only events are dropped (> 32: never)
constexpr uint32_t UN_LUT_BYTES = 256 * 256; // 256-byte row per byte value; lanes use the first 128 B
constexpr uint32_t FULL = 0xffffffffu;

// Dynamic shared memory.  The LUT must start at a 64 KB-aligned shared address; the gap in front of it
// (63 KB when the dynamic window starts at 0x400, the usual case) holds the event lists and the counters;
// the per-warp row rings follow the LUT, the verification tables follow the rings (as long as they fit).
constexpr uint32_t UN_Q_BYTES = UN_WARPS * UN_QCAP * UN_Q_WORDS * 4;
constexpr uint32_t UN_RING_BYTES = UN_WARPS * UN_SLOTS * UN_SLOT_BYTES;
constexpr uint32_t UN_SCRATCH_BYTES = UN_WARPS * 256;
constexpr uint32_t UN_FRONT_FIXED = UN_Q_BYTES + UN_SCRATCH_BYTES + 16;
constexpr uint32_t UN_FRONT_MAX = 60 * 1024; // what the gap is trusted to hold
constexpr size_t UN_SMEM_BYTES = 65536 + UN_LUT_BYTES + UN_RING_BYTES;
static_assert(UN_FRONT_FIXED + 1024 <= UN_FRONT_MAX, "event lists do not fit in front of the LUT");
constexpr size_t UN_SMEM_MAX = 227 * 1024; // opt-in shared memory of one block on sm_100
static_assert(UN_SMEM_BYTES <= UN_SMEM_MAX, "shared memory budget");

struct union_params {
    const uint8_t *bytes; // device pointer to absolute byte abs_base (abs_base % 512 == 0)
    uint64_t abs_base;
    const uint64_t *offsets; // [n_packets+1], absolute
    uint32_t n_packets;
    const uint32_t *items; // [n_items+1] first packet of each work item
    uint32_t n_items;
    uint32_t *work;         // [0] next item, [1] error flags
    const uint32_t *filter; // [256]
    uint32_t n_uniq;
    uint32_t counts_in_smem; // counters live in shared memory
    uint32_t vtab_in_smem;   // the hash verification tables live in shared memory
    const uint32_t *vtab;    // hash verification tables (automaton.c build_verify_tables)
    uint32_t vtab_words;
    uint32_t vtab_one_off;   // vtab[5]: the one-byte patterns' table, 0 = none
    uint32_t mul256; // the value 256 (64 with KMPB_FILTER6), passed at run time so the shift-or-0xff compiles to an
                     // integer multiply-add on the FMA pipe instead of competing for the ALU pipe
    uint32_t mul4096; // KMPB_FILTER6: the two-byte shift
    unsigned long long *uniq_counts;
    // fused expansion + reduction (n_out > 0): the last block to finish adds every pattern's count, in file
    // order, to out[0..n_out) -- this GPU's count vector and/or its peers' (NVLink-mapped)
    const uint32_t *pat_to_uniq;
    uint32_t n_pat, n_out;
    unsigned long long *out[KMPB_MAX_OUT];
};

// ---- work partition: item i = packets [items[i], items[i+1]) ---------------------------------
__global__ void kmpb_union_partition_kernel(const uint64_t *__restrict__ offsets, uint32_t n_packets,
                                            uint32_t n_items, uint32_t n_big, uint64_t item_bytes, uint64_t small_bytes,
                                            uint32_t *__restrict__ items,
                                            uint32_t *__restrict__ work)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { work[0] = 0; work[1] = 0; work[2] = 0; }
    if (i > n_items) return;
    if (i == n_items) { items[i] = n_packets; return; }
    // first packet whose start is >= the item's target byte: n_big items of item_bytes, then small
    // items, so that the warps run dry within a small item of each other at the end of the batch
    const uint64_t target = offsets[0] + (i <= n_big ? (uint64_t)i * item_bytes
                                                     : (uint64_t)n_big * item_bytes + (uint64_t)(i - n_big) * small_bytes);
    // guess by the mean packet size (exact for equal-sized packets), bracket the answer by doubling steps around the
    // guess, then bisect: a few dependent loads instead of log2(n_packets)
    const uint64_t first = offsets[0], total = offsets[n_packets] - first;
    uint32_t lo = 0, hi = n_packets;
    if (total) {
        const double frac = (double)(target - first) / (double)total;
        uint32_t g = (uint32_t)(frac * (double)n_packets);
        g = g > n_packets ? n_packets : g;
        if (offsets[g] < target) { // the answer lies above g
            lo = g + 1;
            for (uint32_t step = 1; lo < n_packets; step *= 2) {
                const uint32_t probe = lo + step - 1 < n_packets ? lo + step - 1 : n_packets;
                if (offsets[probe] >= target) { hi = probe; break; }
                lo = probe + 1;
                if (probe == n_packets) break;
            }
            if (lo > hi) lo = hi;
        } else { // at or below g
            hi = g;
            for (uint32_t step = 1; hi > 0; step *= 2) {
                const uint32_t probe = hi > step ? hi - step : 0;
                if (offsets[probe] < target) { lo = probe + 1; break; }
                hi = probe;
            }
        }
    }
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (offsets[mid] < target) lo = mid + 1; else hi = mid;
    }
    items[i] = lo;
}

// ---- helpers -----------------------------------------------------------------------------------
// per-lane 16-byte asynchronous copies global -> shared (SASS LDGSTS), completion by commit groups
__device__ __forceinline__ void cp_async16(uint32_t dst_sa, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_sa), "l"(src) : "memory");
}
// pointer + 32-bit offset as IMAD.WIDE.U32 (FMA pipe), not IADD3 + IADD3.X (ALU pipe)
__device__ __forceinline__ const uint8_t *add_wide(const uint8_t *base, uint32_t offset)
{
    uint64_t r;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(offset), "l"(reinterpret_cast<uint64_t>(base)));
    return reinterpret_cast<const uint8_t *>(r);
}
// the same under a predicate (no branch around a single copy)
__device__ __forceinline__ void cp_async16_if(bool on, uint32_t dst_sa, const void *src)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p cp.async.cg.shared.global [%1], [%2], 16;\n\t}"
                 ::"r"((uint32_t)on), "r"(dst_sa), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds128v(uint32_t saddr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}
// loads through 32-bit shared addresses (read-only tables, or data ordered by __syncwarp)
__device__ __forceinline__ uint32_t lds32(uint32_t saddr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t lds8v(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t saddr)
{
    uint2 v;
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 lds64v(uint32_t saddr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds32v(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128v(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 4 text bytes starting at byte `pos` of an event (pos + 4 <= 40)
__device__ __forceinline__ uint32_t entry_window(uint32_t entry_sa, uint32_t pos)
{
    const uint32_t a = entry_sa + (pos & ~3u);
    return __funnelshift_r(lds32v(a), lds32v(a + 4), 8u * (pos & 3u));
}
__device__ __forceinline__ uint32_t saddr_of(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// bits [lo, hi) of a 32-bit word, 0 <= lo, hi <= 32
__device__ __forceinline__ uint32_t bit_window(uint32_t lo, uint32_t hi)
{
    const uint32_t below_hi = hi >= 32 ? FULL : (1u << hi) - 1u;
    const uint32_t below_lo = lo >= 32 ? FULL : (1u << lo) - 1u;
    return below_hi & ~below_lo;
}

// filter word of byte `sel` of `word`: the PRMT builds the whole shared address
// [lutlane.b0 | byte | lutlane.b2 | lutlane.b3], lutlane = 64 KB-aligned LUT base + 4*lane
#define LUT_AT(word, sel) lds32(__byte_perm((word), lutlane, (sel)))
#define SEL0 0x7604
#define SEL1 0x7614
#define SEL2 0x7624
#define SEL3 0x7634
#ifdef KMPB_FILTER6
// Experimental (not the default): the filter in 6-bit fields -- 5 pattern buckets + the NUL
// detector, depth 4, and a fifth field in which a report lingers one step (automaton.c kmpb_filter6_build) -- so that
// the row loop updates the state once per TWO bytes: S = ((S << 12) | 0xfff) & G[b0] & L[b1], G = (L << 6) | 0x3f in
// the second half of the LUT's 256-byte rows.  After the update bits 24..29 are b0's reports, bits 18..23 b1's.
#define F6_ARM 0x00020820u                                               // NUL stage pre-armed at depths 0..2
#define LUT_G(word, sel) lds32(__byte_perm((word), lutlane + 128u, (sel)))
#define SA_NEXT(word, sel) (S = (S * mul + 63u) & LUT_AT(word, sel))     // one byte (the resolve step's re-run)
#define SA2_STEP(word, selA, selB, acc)                                        \
    do {                                                                       \
        S = (S * mul2 + 4095u) & LUT_G(word, selA) & LUT_AT(word, selB);       \
        acc |= S;                                                              \
    } while (0)
#define SA2_WORD(word, acc)               \
    do {                                  \
        SA2_STEP(word, SEL0, SEL1, acc);  \
        SA2_STEP(word, SEL2, SEL3, acc);  \
    } while (0)
#else
#define SA_NEXT(word, sel) (S = (S * mul + 255u) & LUT_AT(word, sel))
#endif
#define SA_STEP(word, sel, acc) \
    do {                        \
        SA_NEXT(word, sel);     \
        acc |= S;               \
    } while (0)
#define SA_WORD(word, acc)        \
    do {                          \
        SA_STEP(word, SEL0, acc); \
        SA_STEP(word, SEL1, acc); \
        SA_STEP(word, SEL2, acc); \
        SA_STEP(word, SEL3, acc); \
    } while (0)
// same step, shifting "this byte is NUL" into zr (the first step ends up in the highest bit used) ...
#ifdef KMPB_FILTER6
#define SZ_STEP(word, sel)                   \
    do {                                     \
        SA_NEXT(word, sel);                  \
        zr = __funnelshift_l(S << 8, zr, 1); \
    } while (0)
#define SV_STEP(word, sel)                                              \
    do {                                                                \
        SZ_STEP(word, sel);                                             \
        cmr = __funnelshift_l((S & 0x007c0000u) + 0x7ffc0000u, cmr, 1); \
    } while (0)
#else
#define SZ_STEP(word, sel)                 \
    do {                                   \
        SA_NEXT(word, sel);                \
        zr = __funnelshift_l(S, zr, 1);    \
    } while (0)
// ... and "a candidate start fired" into cmr
#define SV_STEP(word, sel)                                              \
    do {                                                                \
        SZ_STEP(word, sel);                                             \
        cmr = __funnelshift_l((S & 0x7f000000u) + 0x7f000000u, cmr, 1); \
    } while (0)
#endif
#define SV_WORD(word)        \
    do {                     \
        SV_STEP(word, SEL0); \
        SV_STEP(word, SEL1); \
        SV_STEP(word, SEL2); \
        SV_STEP(word, SEL3); \
    } while (0)

// what the slow path needs besides the event list
struct slow_ctx {
    const uint8_t *bytes; // absolute byte abs_base
    uint64_t abs_base;
    const uint64_t *offsets;
    const uint32_t *vtab_g; // verification tables in global memory
    uint32_t vtab_sa;       // ... or their shared address (vtab_in_smem)
    uint32_t vtab_in_smem;
    uint32_t one_off;       // word offset of the one-byte patterns' table, 0 = none
    uint32_t scratch_sa;    // per-warp scratch: {ks, ke, b_abs, e_abs} of the items of either parity (2 x 32 B),
                            // 32 words the resolve step publishes, item parity, the "previous item pending" flag, the NUL carry
    uint32_t s_counts_sa;   // shared address of the shared counters, or 0
    unsigned long long *g_counts;
};

// the hot fields of slow_ctx, read once per resolve step (the context itself lives in local memory)
struct verify_ctx {
    uint32_t vtab_sa, one_off, s_counts_sa;
};

__device__ __forceinline__ void count_hit(const slow_ctx &c, const verify_ctx &v, uint32_t u)
{
    if (v.s_counts_sa) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(v.s_counts_sa + 4u * u) : "memory");
    else atomicAdd(c.g_counts + u, 1ull);
}

// 4 text bytes starting at byte `pos` of the event at entry_sa, of which the first `need` (>= 1) matter: from
// the event while its 40 bytes last, then from global memory (never past the word that holds the last
// needed byte, which lies inside the packet)
__device__ __forceinline__ uint32_t text_window(const slow_ctx &c, uint32_t entry_sa, uint32_t pos, uint32_t need)
{
    if (pos + 4 <= 40) return entry_window(entry_sa, pos);
    const uint8_t *gw = c.bytes + (uint64_t)UN_GRP * (lds32v(entry_sa + 40) & 0x7fffffffu) + (pos & ~3u);
    const uint32_t lo = __ldg(reinterpret_cast<const uint32_t *>(gw));
    const uint32_t hi = (pos & 3u) + (need < 4 ? need : 4u) > 4u ? __ldg(reinterpret_cast<const uint32_t *>(gw) + 1) : 0u;
    return __funnelshift_r(lo, hi, 8u * (pos & 3u));
}

// Every pattern that starts at byte `i` of the event at entry_sa and is at most `room` (>= 1) bytes long
// is counted.  The candidate's first two bytes select one slot of the verification tables (automaton.c
// build_verify_tables); the slot's records -- the patterns whose first two bytes hash there, usually those of one
// two-byte prefix -- carry the pattern's first 8 bytes and their masks, so a record costs one 16-byte load and one
// masked compare, and only a record that agrees on those bytes is looked at further.
template <bool VS>
__device__ __forceinline__ void verify_start(const slow_ctx &c, const verify_ctx &v, const uint4 hdr, uint32_t entry_sa, uint32_t i,
                                             uint32_t room)
{
    auto vt = [&](uint32_t word) -> uint32_t { return VS ? lds32(v.vtab_sa + 4u * word) : __ldg(c.vtab_g + word); };
    auto vt4 = [&](uint32_t word) -> uint4 {
        return VS ? lds128v(v.vtab_sa + 4u * word) : __ldg(reinterpret_cast<const uint4 *>(c.vtab_g + word));
    };
    auto vt2 = [&](uint32_t word) -> uint2 {
        return VS ? lds64(v.vtab_sa + 4u * word) : __ldg(reinterpret_cast<const uint2 *>(c.vtab_g + word));
    };
    // hdr = vtab[1..4]: slots, hash shift, records, pattern words
    const uint32_t x0 = entry_window(entry_sa, i);
    if (v.one_off) { // one-byte patterns: a direct table (room >= 1 always holds)
        const uint32_t u = vt(v.one_off + (x0 & 0xffu));
        if (u != 0xffffffffu) count_hit(c, v, u);
    }
    const uint2 e = vt2(hdr.x + 2u * (((x0 & 0xffffu) * 0x9e3779b1u) >> hdr.y)); // {first record, records}
    if (e.y == 0 || room < 2) return;
    const uint32_t x1 = room > 4 ? text_window(c, entry_sa, i + 4, room - 4) : 0u; // text bytes i+4..i+7
    for (uint32_t r = hdr.z + 8u * e.x, rend = r + 8u * e.y; r != rend; r += 8) {
        const uint4 a = vt4(r); // pattern bytes 0..3, their mask, bytes 4..7, their mask
        if ((((x0 ^ a.x) & a.y) | ((x1 ^ a.z) & a.w)) != 0) continue;
        const uint2 b = vt2(r + 4); // length, distinct id
        const uint32_t m = b.x;
        // m <= room: else it would end past the packet (serial.c:191: the text ends there); the bytes of x0/x1 past
        // `room` are not text of this packet, but a pattern that short compares none of them
        if (m > room) continue;
        bool same = true;
        if (m > 8) {
            const uint32_t pw0 = hdr.w + vt(r + 6);
            for (uint32_t j = 8; j < m && same; j += 4) { // pattern bytes j..j+3 against text bytes i+j..
                const uint32_t pw = vt(pw0 + (j >> 2)), rem = m - j;
                const uint32_t diff = text_window(c, entry_sa, i + j, rem) ^ pw;
                same = (rem >= 4 ? diff : diff & ((1u << (8 * rem)) - 1u)) == 0;
            }
        }
        if (same) count_hit(c, v, b.y);
    }
}

// Resolve the warp's n pending events (n <= 32).  The warp's carry -- 1 + absolute position of the last NUL byte
// seen in the events it has resolved so far (0 = none) -- lives in its scratch words (bytes 200..207), not in a
// register of the row loop.
//
// Phase 1, one event per lane: which start positions fired, where the NULs are, which packet(s) the
// group lies in -> mask of candidate starts that are alive (inside the item, no NUL before them in their
// packet) and mask of packet boundaries inside the group.
// Phase 2, one alive candidate per lane, whichever event it came from: hash lookup and count.
//
//
// The pending events belong to the work item the warp is scanning or to the one before it (the row loop sees to
// that); an event carries its item's parity, and the packets [ks, ke) and bytes [b_abs, e_abs) of both items wait in
// the warp's scratch words, where the warp put them when it took the item: no lane has to look them up in global
// memory, and the row loop does not keep them in registers.
__device__ __noinline__ void drain_events(const slow_ctx &c, const uint32_t q_sa, const uint32_t n, const uint32_t lutlane,
                                          const uint32_t mul)
{
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();
    const uint32_t entry_sa = q_sa + lane * (UN_Q_WORDS * 4);
    uint32_t cm = 0, zm = 0;
    uint64_t gq = 0, b_abs = 0, e_abs = 0; // my group's first byte; my item's byte range (absolute)
    uint32_t ks = 0, ke = 0;
    if (lane < n) {
        const uint2 t = lds64v(entry_sa + 40); // group index | item parity << 31, quarter reports
        gq = c.abs_base + (uint64_t)UN_GRP * (t.x & 0x7fffffffu);
        const uint32_t set_sa = c.scratch_sa + ((t.x >> 31) << 5); // my item's {ks, ke, b_abs, e_abs}
        const uint4 iw = lds128v(set_sa);
        const uint2 iw2 = lds64v(set_sa + 16);
        ks = iw.x;
        ke = iw.y;
        b_abs = (uint64_t)iw.w << 32 | iw.z;
        e_abs = (uint64_t)iw2.y << 32 | iw2.x;
        // Re-run the filter over the quarters that reported, this time recording which start positions
        // fired and which bytes are NUL.  Quarter k: bytes 8k..8k+10 (three bytes of run-in, then starts
        // 8k..8k+7 report at bytes 8k+3..8k+10); the NUL bit is exact from the first byte on.
#ifdef KMPB_UN_LEAN_EVENTS
        // Experimental (not the default, not yet run on a GPU): the row loop pushed only the
        // group index and the reports; the group's 32 bytes and the 8 after them come from L2 now (the row was read
        // a few microseconds ago) and go into the event slot, where the rest of the resolve step expects them.
        // Readable: the batch up to its end rounded up to 32 (include/kmpb200.h); a group at or past the item's end
        // reported from stale ring bytes and holds nothing of this item.
        uint32_t reports = t.y;
        {
            uint4 g0 = make_uint4(0, 0, 0, 0), g1 = g0;
            uint2 g2 = make_uint2(0, 0);
            if (gq < e_abs) {
                const uint8_t *src = c.bytes + (uint64_t)UN_GRP * (t.x & 0x7fffffffu);
                g0 = __ldg(reinterpret_cast<const uint4 *>(src));
                g1 = __ldg(reinterpret_cast<const uint4 *>(src + 16));
                if (gq + UN_GRP < e_abs) g2 = __ldg(reinterpret_cast<const uint2 *>(src + 32));
            } else {
                reports = 0;
            }
            sts128v(entry_sa, g0.x, g0.y, g0.z, g0.w);
            sts128v(entry_sa + 16, g1.x, g1.y, g1.z, g1.w);
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(entry_sa + 32), "r"(g2.x), "r"(g2.y) : "memory");
        }
#else
        const uint32_t reports = t.y;
#endif
        uint32_t quarters = ((((reports & 0x7f7f7f7fu) + 0x7f7f7f7fu) | reports) & 0x80808080u); // bit 8k+7: quarter k
        while (quarters) {
            const uint32_t k8 = (__ffs(quarters) - 1) & ~7u; // 8k
            quarters &= quarters - 1;
            const uint32_t w0 = lds32v(entry_sa + k8), w1 = lds32v(entry_sa + k8 + 4), w2 = lds32v(entry_sa + k8 + 8);
#ifdef KMPB_FILTER6
            // quarter k here: bytes 8k..8k+11, starts 8k..8k+8 (the quarters of the row loop overlap by one start;
            // OR-ing a start bit twice changes nothing)
            uint32_t S = F6_ARM, cmr = 0, zr = 0;
            SZ_STEP(w0, SEL0); SZ_STEP(w0, SEL1); SZ_STEP(w0, SEL2);
            SV_STEP(w0, SEL3);
            SV_WORD(w1);
            SV_WORD(w2);
            cm |= (__brev(cmr) >> 23) << k8;
            zm |= (__brev(zr) >> 20) << k8;
#else
            uint32_t S, cmr = 0, zr;
            S = LUT_AT(w0, SEL0) & 0x808080ffu;
            zr = S >> 31;
            SZ_STEP(w0, SEL1); SZ_STEP(w0, SEL2);
            SV_STEP(w0, SEL3);
            SV_WORD(w1);
            SV_STEP(w2, SEL0); SV_STEP(w2, SEL1); SV_STEP(w2, SEL2);
            cm |= (__brev(cmr) >> 24) << k8;
            zm |= (__brev(zr) >> 21) << k8;
#endif
        }
        // the item's first and last rows overhang it: starts count inside [b_abs, e_abs) only, and bytes
        // past e_abs may be stale ring contents, so NULs count below e_abs only
        const uint32_t lo = b_abs > gq ? (uint32_t)min(b_abs - gq, (uint64_t)32) : 0u;
        const uint32_t hi = e_abs > gq ? (uint32_t)min(e_abs - gq, (uint64_t)32) : 0u;
        cm &= bit_window(lo, hi);
        zm &= bit_window(0, hi);
    }
    // last NUL before my group: the nearest earlier event that holds one, else the warp's carry
    const uint64_t mylast1 = zm ? gq + (32u - __clz(zm)) : 0ull; // 1 + position of my last NUL
    const uint32_t nulm = __ballot_sync(FULL, zm != 0);
    const uint32_t below = nulm & ((1u << lane) - 1u);
    const uint64_t from_below = __shfl_sync(FULL, mylast1, below ? 31 - __clz(below) : 0);
    const uint2 cw = lds64v(c.scratch_sa + 200);
    const uint64_t carry = (uint64_t)cw.y << 32 | cw.x;
    const uint64_t prev1 = below ? from_below : carry;
    const uint64_t from_top = __shfl_sync(FULL, mylast1, nulm ? 31 - __clz(nulm) : 0);
    __syncwarp();
    if (nulm && lane == 0)
        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(c.scratch_sa + 200), "r"((uint32_t)from_top), "r"((uint32_t)(from_top >> 32)) : "memory");

    uint32_t am = 0, bm = 0, nextb = 255; // alive candidates; packet starts inside the group (bit = offset);
                                          // offset of the first packet start at or after the group's end
    if (cm) {
        // The packet that holds my first candidate: the last k in [ks, ke) with offsets[k] <= p0.  First
        // guess by interpolation (exact for equal-sized packets), then a binary search in what is left.
        const uint64_t p0 = gq + (__ffs(cm) - 1);
        uint32_t k = ks, k1 = ke;
        {
            const float frac = (float)(uint32_t)(p0 - b_abs) / (float)(uint32_t)(e_abs - b_abs);
            uint32_t kg = ks + (uint32_t)(frac * (float)(ke - ks));
            kg = kg >= ke ? ke - 1 : kg;
            const uint64_t o0 = __ldg(c.offsets + kg), o1 = __ldg(c.offsets + kg + 1);
            if (o0 <= p0) {
                k = kg;
                if (p0 < o1) k1 = kg + 1;
                else k = kg + 1; // offsets[kg + 1] <= p0 < e_abs, so kg + 1 < ke
            } else {
                k1 = kg;
            }
        }
        while (k1 - k > 1) {
            const uint32_t mid = k + (k1 - k) / 2;
            if (__ldg(c.offsets + mid) <= p0) k = mid; else k1 = mid;
        }
        uint64_t ps = __ldg(c.offsets + k), pe = __ldg(c.offsets + k + 1);
        bool dead = prev1 > ps; // a NUL in [ps, group): kmp_matcher's strlen() stopped before my group
        uint32_t a = ps > gq ? (uint32_t)(ps - gq) : 0u; // the packet [ps, pe) covers my group from offset a on
        for (;;) {
            const uint32_t b = pe - gq >= 32 ? 32u : (uint32_t)(pe - gq);
            const uint32_t seg = bit_window(a, b), z = zm & seg;
            if (!dead) am |= cm & seg & (z ? (z & (0u - z)) - 1u : FULL); // starts before the packet's first NUL
            if (b == 32) break;
            bm |= 1u << b;          // the next packet starts inside my group
            if (k + 1 >= ke) break; // ... or the item ends there: nothing beyond is mine
            a = b;
            dead = false;
            k++;
            ps = pe;
            pe = __ldg(c.offsets + k + 1);
        }
        nextb = pe - gq > 255 ? 255u : (uint32_t)(pe - gq);
    }
    // publish what phase 2 needs next to the event's bytes
    if (lane < n) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(entry_sa + 44), "r"(bm) : "memory"); // over the quarter reports
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.scratch_sa + 64 + 4 * lane), "r"(nextb) : "memory");
    }
    // alive candidates, numbered across the lanes
    const uint32_t cnt = __popc(am);
    uint32_t incl = cnt;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    const uint32_t excl = incl - cnt;
    __syncwarp();
    verify_ctx vc;
    vc.vtab_sa = c.vtab_sa;
    vc.one_off = c.one_off;
    vc.s_counts_sa = c.s_counts_sa;
    // vtab[1..4]: where the slots, the records and the pattern words are
    const uint4 vhdr = total == 0 ? make_uint4(0, 0, 0, 0)
                       : c.vtab_in_smem ? make_uint4(lds32(c.vtab_sa + 4), lds32(c.vtab_sa + 8), lds32(c.vtab_sa + 12), lds32(c.vtab_sa + 16))
                                        : make_uint4(__ldg(c.vtab_g + 1), __ldg(c.vtab_g + 2), __ldg(c.vtab_g + 3), __ldg(c.vtab_g + 4));
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
        const uint32_t t = t0 + lane;
        // owner of candidate t: the first lane whose inclusive count exceeds t
        uint32_t l = 0;
#pragma unroll
        for (uint32_t s = 16; s; s >>= 1) {
            const uint32_t v = __shfl_sync(FULL, incl, l + s - 1);
            if (v <= t) l += s;
        }
        l &= 31;
        uint32_t m = __shfl_sync(FULL, am, l);
        const uint32_t first = __shfl_sync(FULL, excl, l);
        if (t < total) {
            for (uint32_t j = first; j < t; j++) m &= m - 1;
            const uint32_t i = __ffs(m) - 1;
            const uint32_t owner_sa = q_sa + l * (UN_Q_WORDS * 4);
            const uint32_t obm = lds32v(owner_sa + 44), onext = lds32v(c.scratch_sa + 64 + 4 * l);
            const uint32_t above = obm & ~((2u << i) - 1u); // packet starts after byte i
            const uint32_t room = (above ? (uint32_t)__ffs(above) - 1u : onext) - i;
            if (c.vtab_in_smem) verify_start<true>(c, vc, vhdr, owner_sa, i, room);
            else verify_start<false>(c, vc, vhdr, owner_sa, i, room);
        }
    }
    __syncwarp();
}

// NUL-dense row (many groups reported and the list is full).  An event without candidates exists only to tell later
// candidates where the last NUL before them is; when the next event of the row is of the same kind, that one tells
// them a later NUL and this one is not needed: its reports are cleared (binary payloads: one event per row instead
// of one per group).  m = lanes with reports.  Kept out of line: the row loop should stay small.
__device__ __noinline__ uint32_t drop_superseded(uint32_t tops, const uint32_t m)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t mc = __ballot_sync(FULL, (tops & 0x7f7f7f7fu) != 0); // lanes with candidates
    const uint32_t above = m & ~((2u << lane) - 1u);                    // reporting lanes above me
    if (!((mc >> lane) & 1u) && above && !((mc >> (__ffs(above) - 1)) & 1u)) tops = 0;
    return tops;
}

// warp-uniform value, in a form the compiler can keep in a uniform register
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __shfl_sync(FULL, v, 0); }
__device__ __forceinline__ uint64_t uni(uint64_t v) { return __shfl_sync(FULL, v, 0); }
__global__ void __launch_bounds__(UN_THREADS, 1) kmpb_union_kernel(const __grid_constant__ union_params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // the LUT sits at the first 64 KB-aligned shared address inside the dynamic allocation
    const uint32_t dyn_saddr = saddr_of(smem);
    uint32_t dyn_size;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    const uint32_t lut_off = (0x10000u - (dyn_saddr & 0xffffu)) & 0xffffu;
    const uint32_t counts_bytes = p.counts_in_smem ? ((4u * p.n_uniq + 15u) & ~15u) : 0u;
    const uint32_t vtab_bytes = p.vtab_in_smem ? 4u * p.vtab_words : 0u;
    const uint32_t front_bytes = UN_FRONT_FIXED + counts_bytes;
    if (lut_off < front_bytes || lut_off + UN_LUT_BYTES + UN_RING_BYTES + vtab_bytes > dyn_size) {
        // unexpected shared-memory window base: refuse rather than compute something wrong
        if (threadIdx.x == 0) atomicOr(&p.work[1], 2u);
        return;
    }
    uint8_t *lut = smem + lut_off;
    uint8_t *ring_all = lut + UN_LUT_BYTES;
    uint8_t *q_all = smem;
    uint8_t *scratch_all = smem + UN_Q_BYTES;
    uint32_t *s_lut_saddr = reinterpret_cast<uint32_t *>(scratch_all + UN_SCRATCH_BYTES);
    uint32_t *s_counts = reinterpret_cast<uint32_t *>(s_lut_saddr + 4);
    uint32_t *s_vtab = reinterpret_cast<uint32_t *>(ring_all + UN_RING_BYTES); // behind the rings

    for (uint32_t i = threadIdx.x; i < 256 * 32; i += UN_THREADS)
#ifdef KMPB_FILTER6
    {
        const uint32_t f = p.filter[i >> 5];
        reinterpret_cast<uint32_t *>(lut)[(i >> 5) * 64 + (i & 31)] = f;
        reinterpret_cast<uint32_t *>(lut)[(i >> 5) * 64 + 32 + (i & 31)] = (f << 6) | 0x3fu;
    }
#else
        reinterpret_cast<uint32_t *>(lut)[(i >> 5) * 64 + (i & 31)] = p.filter[i >> 5];
#endif
    if (p.counts_in_smem)
        for (uint32_t i = threadIdx.x; i < p.n_uniq; i += UN_THREADS) s_counts[i] = 0;
    if (p.vtab_in_smem)
        for (uint32_t i = threadIdx.x; i < p.vtab_words; i += UN_THREADS) s_vtab[i] = p.vtab[i];
    if (threadIdx.x == 0) *s_lut_saddr = dyn_saddr + lut_off;
    if (threadIdx.x < UN_WARPS) { // per-warp item parity and "previous item pending" flag
        reinterpret_cast<uint32_t *>(scratch_all + threadIdx.x * 256)[48] = 0;
        reinterpret_cast<uint32_t *>(scratch_all + threadIdx.x * 256)[49] = 0;
        reinterpret_cast<uint32_t *>(scratch_all + threadIdx.x * 256)[50] = 0; // the warp's NUL carry
        reinterpret_cast<uint32_t *>(scratch_all + threadIdx.x * 256)[51] = 0;
    }
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = uni(threadIdx.x >> 5);
    // per-warp row ring: UN_SLOTS slots of one row (+16 bytes) each
    // Inside a slot lane l owns bytes [32 l, 32 l + 32); lanes 4..7, 12..15, ... keep their two 16-byte
    // halves swapped, which makes the 16-byte copies and reads of a quarter-warp hit 8 different bank groups.
    const uint32_t ring_sa = saddr_of(ring_all) + warp * (UN_SLOTS * UN_SLOT_BYTES);
    const uint32_t offtail = UN_ROW;
#ifdef KMPB_UN_NOSWIZZLE
    const uint32_t off0 = lane * UN_GRP, off1 = off0 + 16, offla = off0 + UN_GRP;
#else
    const uint32_t off0 = lane * UN_GRP + ((lane >> 2) & 1u) * 16, off1 = off0 ^ 16u;
#ifdef KMPB_UN_COALESCED_COPY
    const uint32_t coff = (lane * 16) ^ (((lane >> 3) & 1u) << 4); // slot offset of row bytes [16 lane, 16 lane + 16)
#endif
    // lane 31's lookahead: the slot's tail, or (three slots) the first bytes of the following slot
    const uint32_t offla = lane == 31 ? UN_ROW : (lane + 1) * UN_GRP + (((lane + 1) >> 2) & 1u) * 16;
#endif
    __syncthreads();

    // read back through shared memory so that no LUT load can be scheduled above the barrier
    const uint32_t lutlane = *s_lut_saddr + (lane << 2);
    const uint32_t mul = p.mul256;
#ifdef KMPB_FILTER6
    const uint32_t mul2 = p.mul4096;
#endif
    uint32_t lt;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt));
    const uint32_t q_sa = saddr_of(q_all) + warp * (UN_QCAP * UN_Q_WORDS * 4);
    slow_ctx sc;
    sc.bytes = p.bytes;
    sc.abs_base = p.abs_base;
    sc.offsets = p.offsets;
    sc.vtab_g = p.vtab;
    sc.vtab_sa = saddr_of(s_vtab);
    sc.vtab_in_smem = p.vtab_in_smem;
    sc.one_off = p.vtab_one_off;
    sc.scratch_sa = saddr_of(scratch_all) + warp * 256;
    sc.s_counts_sa = p.counts_in_smem ? saddr_of(s_counts) : 0u;
    sc.g_counts = p.uniq_counts;

    uint32_t qn = 0;         // pending events

    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&p.work[0], 1u);
        item = uni(item);
        if (item >= p.n_items) break;
        const uint32_t ks = uni(p.items[item]), ke = uni(p.items[item + 1]);
        if (ks >= ke) continue;
        const uint64_t b_abs = uni(p.offsets[ks]), e_abs = uni(p.offsets[ke]);
        if (b_abs == e_abs) continue;
        if (e_abs - b_abs >= (1ull << 31)) { // a packet over 2 GiB: outside the documented limits
            if (lane == 0) atomicOr(&p.work[1], 1u);
            continue;
        }
        // What the slow path needs to know about this item goes into the scratch set of the item's parity.  The set
        // still belongs to the item before the previous one; its events are gone unless the list has not been
        // emptied since then (flag word 49: "the list holds events of the previous item"), in which case it is now.
        const uint2 st = lds64v(sc.scratch_sa + 192); // parity of the previous item, pending flag
        const uint32_t par = st.x ^ 1u;
        if (qn && st.y) {
            drain_events(sc, q_sa, qn, lutlane, mul);
            qn = 0;
        }
        __syncwarp();
        if (lane == 0) {
            const uint32_t set_sa = sc.scratch_sa + (par << 5);
            sts128v(set_sa, ks, ke, (uint32_t)b_abs, (uint32_t)(b_abs >> 32));
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(set_sa + 16), "r"((uint32_t)e_abs), "r"((uint32_t)(e_abs >> 32)) : "memory");
            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(sc.scratch_sa + 192), "r"(par), "r"(qn ? 1u : 0u) : "memory");
        }
        const uint64_t row0 = b_abs & ~127ull; // absolute position of the item's first row
#ifdef KMPB_UN_COALESCED_COPY
        const uint8_t *textl = p.bytes + (row0 - p.abs_base) + lane * 16; // the 16 bytes of the first row I copy first
#else
        const uint8_t *textl = p.bytes + (row0 - p.abs_base) + lane * UN_GRP; // my 32 bytes of the item's first row
#endif
        const uint32_t e_rel = (uint32_t)(e_abs - row0);
        const uint32_t load_end = (e_rel + 15u) & ~15u;
        const uint32_t nrows = (e_rel + UN_ROW - 1) / UN_ROW;
        // my group's index in row 0, in 32-byte units (< 2^31: a batch is below 64 GiB), and the item's parity
        const uint32_t g32 = ((uint32_t)((row0 - p.abs_base) >> 5) + lane) | par << 31;

        // Every lane copies its own 32 bytes (lane 31 also the 16 bytes after the row -- they follow its own
        // -- into the slot's tail) with 16-byte asynchronous copies.  One commit group per call, also when there is nothing left to copy, so
        // that "all but the newest UN_SLOTS-1 groups are complete" always means "the row about to be
        // scanned has arrived".  (A chunk-major slot layout, free of bank conflicts on both sides, measured
        // 14 % slower on the fast path alone.)
#ifdef KMPB_UN_COALESCED_COPY
        // Which lane copies which 16 bytes is free (the slot is read after a warp barrier): each copy instruction
        // moves 512 CONTIGUOUS bytes -- lane l the bytes [16 l, 16 l + 16) of the row's first and of its second half --
        // instead of every other 16-byte piece.  The per-instruction counters of the capture in profiles/ show why: the
        // strided form costs 14 shared-memory wavefronts per copy instruction instead of 4 (the data arrives by
        // 32-byte sectors of which half is used) and fetches every sector of the row twice from L2.  The byte at row
        // offset x still lands at slot offset x ^ (((x >> 7) & 1) << 4), so nothing else changes.
        auto issue_row = [&](uint32_t r, uint32_t slot) {
            const uint32_t row = r * UN_ROW;
            const uint32_t dst = ring_sa + slot * UN_SLOT_BYTES;
            const uint8_t *src = add_wide(textl, row); // textl: 16 bytes per lane in this form
            if (row + UN_SLOT_BYTES <= load_end) {
                cp_async16(dst + coff, src);
                cp_async16(dst + coff + 512, src + 512);
                if (UN_TAIL) cp_async16_if(lane == 31, dst + offtail, src + (UN_ROW - 16 * 31));
            } else if (row < e_rel) { // r < nrows
                const uint32_t x = row + lane * 16;
                if (x < load_end) cp_async16(dst + coff, src);
                if (x + 512 < load_end) cp_async16(dst + coff + 512, src + 512);
                if (UN_TAIL && lane == 31 && row + UN_ROW < load_end) cp_async16(dst + offtail, src + (UN_ROW - 16 * 31));
            }
            cp_async_commit();
        };
#else
        auto issue_row = [&](uint32_t r, uint32_t slot) {
            const uint32_t row = r * UN_ROW;
            const uint32_t dst = ring_sa + slot * UN_SLOT_BYTES;
            const uint8_t *src = add_wide(textl, row); // one multiply-add on the FMA pipe instead of two ALU adds
            if (row + UN_SLOT_BYTES <= load_end) {
                cp_async16(dst + off0, src);
                cp_async16(dst + off1, src + 16);
                if (UN_TAIL) cp_async16_if(lane == 31, dst + offtail, src + 32);
            } else if (row < e_rel) { // r < nrows
                const uint32_t g = row + lane * UN_GRP;
                if (g < load_end) cp_async16(dst + off0, src);
                if (g + 16 < load_end) cp_async16(dst + off1, src + 16);
                if (UN_TAIL && lane == 31 && g + 32 < load_end) cp_async16(dst + offtail, src + 32);
            }
            cp_async_commit();
        };
#endif
        for (uint32_t r = 0; r < UN_SLOTS; r++) issue_row(r, r);

        // one row: wait for its slot, filter, refill the slot, push the events
        auto scan_row = [&](const uint32_t r, const uint32_t slot) {
            cp_async_wait<UN_TAIL ? UN_SLOTS - 1 : UN_SLOTS - 2>();
            __syncwarp(); // my lookahead is the next lane's copy
            const uint32_t base = ring_sa + slot * UN_SLOT_BYTES;
            const uint4 c0 = lds128v(base + off0), c1 = lds128v(base + off1);
            // the 8 bytes after my group: the next lane's, for lane 31 the first of the next row (next slot)
            const uint2 la2 = lds64v(UN_TAIL || slot + 1 < UN_SLOTS || lane != 31 ? base + offla : ring_sa);
            const uint32_t la = la2.x;

#ifdef KMPB_FILTER6
            // ---- shift-and filter over 36 bytes, two per update -----------------------------------
            // Update j takes bytes 2j, 2j+1 and reports the windows that end there, i.e. the starts 2j-3 and 2j-2.
            // acc[k] collects the starts 0..8, 9..16, 17..24, 25..31 (k = 0 also the NUL bits of bytes 0..2).
            uint32_t S = F6_ARM, acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
            SA2_WORD(c0.x, acc0); SA2_WORD(c0.y, acc0); SA2_WORD(c0.z, acc0); // bytes 0..11
            SA2_WORD(c0.w, acc1); SA2_WORD(c1.x, acc1);                       // bytes 12..19
            SA2_WORD(c1.y, acc2); SA2_WORD(c1.z, acc2);                       // bytes 20..27
            SA2_WORD(c1.w, acc3);                                             // bytes 28..31
            // lookahead: starts 29..31 report here; a NUL here is the next lane's, and so is start 32
            uint32_t accA = 0, accB = 0;
            SA2_STEP(la, SEL0, SEL1, accA); // bytes 32, 33: starts 29, 30
            SA2_STEP(la, SEL2, SEL3, accB); // bytes 34, 35: start 31 (bits 24..28) and start 32
            acc3 |= (accA & 0x1f7c0000u) | (accB & 0x1f000000u);
            // a quarter's reports of either byte of an update, merged into bits 18..23; byte 2 of that is
            // (reports << 2): candidate buckets in bits 2..6, NUL in bit 7 -- the layout the rest of the kernel knows
            acc0 |= acc0 >> 6; acc1 |= acc1 >> 6; acc2 |= acc2 >> 6; acc3 |= acc3 >> 6;
            const uint32_t tops =
                __byte_perm(__byte_perm(acc0, acc1, 0x0062), __byte_perm(acc2, acc3, 0x0062), 0x5410) & 0xfcfcfcfcu;
#else
            // ---- shift-and filter over 35 bytes ---------------------------------------------------
            // Reports are collected per quarter of the group: acc[k] covers the steps at which starts
            // 8k..8k+7 report (and, for k = 0, the three steps before them, for their NUL bits).
            uint32_t S, acc0, acc1, acc2, acc3, accC;
            S = LUT_AT(c0.x, SEL0) & 0x808080ffu; // no history: only the NUL stage is pre-armed
            acc0 = S;
            SA_STEP(c0.x, SEL1, acc0);
            SA_STEP(c0.x, SEL2, acc0);
            SA_STEP(c0.x, SEL3, acc0);
            SA_WORD(c0.y, acc0);
            SA_STEP(c0.z, SEL0, acc0);
            SA_STEP(c0.z, SEL1, acc0);
            SA_STEP(c0.z, SEL2, acc0);
            acc1 = 0;
            SA_STEP(c0.z, SEL3, acc1);
            SA_WORD(c0.w, acc1);
            SA_STEP(c1.x, SEL0, acc1);
            SA_STEP(c1.x, SEL1, acc1);
            SA_STEP(c1.x, SEL2, acc1);
            acc2 = 0;
            SA_STEP(c1.x, SEL3, acc2);
            SA_WORD(c1.y, acc2);
            SA_STEP(c1.z, SEL0, acc2);
            SA_STEP(c1.z, SEL1, acc2);
            SA_STEP(c1.z, SEL2, acc2);
            acc3 = 0;
            SA_STEP(c1.z, SEL3, acc3);
            SA_WORD(c1.w, acc3);
            // lookahead: candidate starts 29..31 report here; a NUL here is the next lane's
            accC = 0;
            SA_STEP(la, SEL0, accC);
            SA_STEP(la, SEL1, accC);
            SA_STEP(la, SEL2, accC);
            acc3 |= accC & 0x7f000000u;
            // the four top bytes side by side: quarter k has something to resolve iff byte k is nonzero
            const uint32_t tops = __byte_perm(__byte_perm(acc0, acc1, 0x0073), __byte_perm(acc2, acc3, 0x0073), 0x5410);
#endif
            const uint32_t m = __ballot_sync(FULL, tops != 0);

            // every lane holds its bytes: refill the slot with the row UN_SLOTS ahead
            issue_row(r + UN_SLOTS, slot);

#ifdef KMPB_ABLATE_SLOW_PATH // measurement only (wrong counts): how fast is the fast path alone?
            if (m == 0x12345678u)
#endif
            if (m) {
                const uint32_t n = __popc(m);
                auto resolve_pending = [&]() {
                    drain_events(sc, q_sa, qn, lutlane, mul);
                    qn = 0;
                    if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sc.scratch_sa + 196), "r"(0u) : "memory"); // nothing older pending
                };
                if (qn + n > UN_QCAP) { // the list cannot take this row's events
                    if (UN_DENSE <= 32 && n >= UN_DENSE) {
                        // NUL-dense row: its own (rare) path, so that the usual one keeps its registers
                        const uint32_t tops2 = drop_superseded(tops, m);
                        const uint32_t m2 = __ballot_sync(FULL, tops2 != 0), n2 = __popc(m2);
                        if (qn + n2 > UN_QCAP) resolve_pending();
                        if (tops2 != 0) {
                            const uint32_t e = q_sa + (qn + __popc(m2 & lt)) * (UN_Q_WORDS * 4);
#ifdef KMPB_UN_LEAN_EVENTS
                            asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(e + 40), "r"(g32 + (r << 5)), "r"(tops2) : "memory");
#else
                            sts128v(e, c0.x, c0.y, c0.z, c0.w);
                            sts128v(e + 16, c1.x, c1.y, c1.z, c1.w);
                            sts128v(e + 32, la, la2.y, g32 + (r << 5), tops2);
#endif
                        }
                        qn += n2;
                        return;
                    }
                    resolve_pending();
                }
                if (tops != 0) {
                    const uint32_t e = q_sa + (qn + __popc(m & lt)) * (UN_Q_WORDS * 4);
#ifdef KMPB_UN_LEAN_EVENTS
                    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(e + 40), "r"(g32 + (r << 5)), "r"(tops) : "memory");
#else
                    sts128v(e, c0.x, c0.y, c0.z, c0.w);
                    sts128v(e + 16, c1.x, c1.y, c1.z, c1.w);
                    sts128v(e + 32, la, la2.y, g32 + (r << 5), tops);
#endif
                }
                qn += n;
            }
        };
#if KMPB_UN_SLOTS == 2
        // two rows per trip: the slots are compile-time constants, half the loop bookkeeping
#pragma unroll 1
        for (uint32_t r = 0; r < nrows; r += 2) {
            scan_row(r, 0);
            if (r + 1 < nrows) scan_row(r + 1, 1);
        }
#elif KMPB_UN_SLOTS == 3
#pragma unroll 1
        for (uint32_t r = 0; r < nrows; r += 3) {
            scan_row(r, 0);
            if (r + 1 < nrows) scan_row(r + 1, 1);
            if (r + 2 < nrows) scan_row(r + 2, 2);
        }
#else
        uint32_t slot = 0;
#pragma unroll 1
        for (uint32_t r = 0; r < nrows; r++) {
            scan_row(r, slot);
            slot = slot + 1 == UN_SLOTS ? 0 : slot + 1;
        }
#endif
    }
    // leftovers
    if (qn) drain_events(sc, q_sa, qn, lutlane, mul);

    __syncthreads();
    if (p.counts_in_smem)
        for (uint32_t u = threadIdx.x; u < p.n_uniq; u += UN_THREADS)
            if (s_counts[u]) atomicAdd(p.uniq_counts + u, (unsigned long long)s_counts[u]);

    // The merge of the partial counts (openmp_data.c:169-173) and, across GPUs, the MPI_Reduce(SUM) of
    // mpi_dumping.c:202, inside this kernel: the last block to arrive expands the distinct-pattern totals
    // to file order and adds them to every count vector it was given -- system-scope atomics, so a vector
    // may live in a peer GPU's memory.
    if (p.n_out) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_lut_saddr[1] = atomicAdd(&p.work[2], 1u) == gridDim.x - 1 ? 1u : 0u;
        __syncthreads();
        if (s_lut_saddr[1]) {
            __threadfence();
            for (uint32_t i = threadIdx.x; i < p.n_pat; i += UN_THREADS) {
                const unsigned long long v = __ldcg(p.uniq_counts + p.pat_to_uniq[i]);
                if (v)
                    for (uint32_t r = 0; r < p.n_out; r++) atomicAdd_system(p.out[r] + i, v);
            }
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------

// scratch for batches of up to max_batch_bytes: work counters and the item table, one set per slot
int kmpb_union_scratch(kmpb_ctx *ctx, uint64_t max_batch_bytes)
{
    size_t need = (size_t)(max_batch_bytes / UN_ITEM_BYTES) + UN_TAIL_ITEMS + 8;
    if (need <= ctx->items_cap && ctx->d_items && ctx->d_work) return KMPB_OK;
    cudaFree(ctx->d_items);
    ctx->d_items = nullptr;
    ctx->items_cap = 0;
    if (!ctx->d_work) KMPB_CUDA(cudaMalloc((void **)&ctx->d_work, KMPB_COPY_STREAMS * 4 * sizeof(uint32_t)));
    KMPB_CUDA(cudaMalloc((void **)&ctx->d_items, (size_t)KMPB_COPY_STREAMS * need * sizeof(uint32_t)));
    ctx->items_cap = need;
    return KMPB_OK;
}

int kmpb_launch_union(kmpb_ctx *ctx, const kmpb_batch &b, int slot, uint64_t *d_uniq_counts, cudaStream_t stream,
                      const kmpb_fused_out &out)
{
    const kmpb_tables &h = ctx->host;
    if (h.n_uniq == 0 || b.n_packets == 0 || b.end_byte == b.first_byte) return KMPB_OK;
    if (b.n_packets >= (1ull << 31)) return kmpb_fail(KMPB_ELIMIT, "more than 2^31-1 packets in one batch");
    if ((b.abs_base & 511) || ((uintptr_t)b.d_bytes & 31))
        return kmpb_fail(KMPB_EINVAL, "payload buffer must be 32-byte aligned");
    if (b.end_byte - b.abs_base >= (1ull << 36))
        return kmpb_fail(KMPB_ELIMIT, "more than 64 GiB of payload in one batch");
    const uint64_t span = b.end_byte - b.first_byte;
    // Taking an item costs a warp three dependent trips to L2 (ticket, item table, offsets); large batches
    // afford larger items (C3, 14 GB: 64 KB items 2.96 TB/s, 256 KB items 3.00 TB/s).
    const uint64_t item_bytes = span >= (8ull << 30) ? 4ull * UN_ITEM_BYTES : span >= (4ull << 30) ? 2ull * UN_ITEM_BYTES : UN_ITEM_BYTES;
    // full-size items, except the last UN_TAIL_ITEMS small items (at most an eighth of the batch), so that the warps
    // run dry within a few rows of each other at the end of the batch: an eighth of a full item, at least 16 KB
    // (C3 with 256 KB items: 64 KB tails 3003 GB/s, 32 KB 3013, 16 KB 2995, 8 KB 2985)
    const uint64_t tail_div = std::min<uint64_t>(8, item_bytes / (UN_ITEM_BYTES / 4));
    const uint64_t whole = span / item_bytes;
    const uint64_t small_bytes = item_bytes / tail_div;
    const uint64_t tail_items = std::min<uint64_t>(UN_TAIL_ITEMS / tail_div, whole / 8); // at most an eighth of the batch
    const uint32_t n_big = (uint32_t)(whole - tail_items);
    const uint64_t small_span = span - (uint64_t)n_big * item_bytes;
    const uint32_t n_items = n_big + (uint32_t)((small_span + small_bytes - 1) / small_bytes);
    if ((size_t)n_items + 1 > ctx->items_cap) return kmpb_fail(KMPB_ESTATE, "union scratch too small");
    uint32_t *d_items = ctx->d_items + (size_t)slot * ctx->items_cap;
    uint32_t *d_work = ctx->d_work + slot * 4;

    // the counters go into the shared-memory gap in front of the LUT if they fit, the hash tables behind
    // the row rings if the block's 227 KB allow it (the LUT may start up to 64 KB into the allocation)
    uint32_t front = UN_FRONT_FIXED;
    const bool counts_in_smem = front + 4ull * h.n_uniq + 16 <= UN_FRONT_MAX;
    if (counts_in_smem) front += (4u * h.n_uniq + 15u) & ~15u;
    const bool vtab_in_smem = UN_SMEM_BYTES + 4ull * h.vtab_words <= UN_SMEM_MAX;
    const size_t smem_bytes = UN_SMEM_BYTES + (vtab_in_smem ? 4ull * h.vtab_words : 0);
    if (!ctx->attr_union_set) {
        KMPB_CUDA(cudaFuncSetAttribute(kmpb_union_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UN_SMEM_MAX));
        ctx->attr_union_set = true;
    }
    kmpb_union_partition_kernel<<<(n_items + 1 + 255) / 256, 256, 0, stream>>>(b.d_offsets, (uint32_t)b.n_packets, n_items,
                                                                              n_big, item_bytes, small_bytes, d_items, d_work);
    union_params p;
    p.bytes = b.d_bytes;
    p.abs_base = b.abs_base;
    p.offsets = b.d_offsets;
    p.n_packets = (uint32_t)b.n_packets;
    p.items = d_items;
    p.n_items = n_items;
    p.work = d_work;
    p.filter = ctx->dev.filter;
    p.n_uniq = h.n_uniq;
    p.counts_in_smem = counts_in_smem ? 1u : 0u;
    p.vtab_in_smem = vtab_in_smem ? 1u : 0u;
    p.vtab = ctx->dev.vtab;
    p.vtab_words = h.vtab_words;
    p.vtab_one_off = h.vtab[5];
#ifdef KMPB_FILTER6
    p.mul256 = 64u;
    p.mul4096 = 4096u;
#else
    p.mul256 = 256u;
    p.mul4096 = 0u;
#endif
    p.uniq_counts = (unsigned long long *)d_uniq_counts;
    p.pat_to_uniq = ctx->dev.pat_to_uniq;
    p.n_pat = h.n_pat;
    p.n_out = out.n;
    for (int r = 0; r < KMPB_MAX_OUT; r++) p.out[r] = out.vec[r];
    const uint32_t warps_needed = n_items;
    int grid = (int)std::min<uint32_t>((uint32_t)ctx->sm_count, (warps_needed + UN_WARPS - 1) / UN_WARPS);
    if (grid < 1) grid = 1;
    if (ctx->profile) KMPB_CUDA(cudaEventRecord(ctx->ev_kernel[0], stream));
    kmpb_union_kernel<<<grid, UN_THREADS, smem_bytes, stream>>>(p);
    if (ctx->profile) KMPB_CUDA(cudaEventRecord(ctx->ev_kernel[1], stream));
    ctx->launches += 2;
    KMPB_CUDA(cudaGetLastError());
    return KMPB_OK;
}
