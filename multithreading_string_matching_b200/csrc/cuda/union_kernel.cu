// union_kernel.cu -- the union engine (KMPB_ENGINE_UNION): every payload byte is read from HBM once.
//
// Two levels.
//
//  FAST PATH (every byte).  A warp streams work items -- runs of whole packets, ~64 KB of the flat
//  CSR byte buffer -- in rows of 1024 contiguous bytes: one coalesced 32-byte load (LDG.256) per lane,
//  two rows in flight ahead of the one being scanned.  Each lane pushes its 32 bytes (+3 bytes of
//  lookahead from its neighbour, by shuffle) through a 4-byte-deep shift-and filter over 8 buckets:
//      S = ((S << 8) | 0xff) & filter[byte]
//  filter[] lives in shared memory in a bank-private layout (byte address = byte*256 + lane*4) at a
//  64 KB-aligned shared address, so the one lookup per byte never bank-conflicts and its complete
//  address is a single PRMT of the text word with a per-lane constant.  The shift-or-0xff is one
//  integer multiply-add (FMA pipe), the AND one LOP3 (ALU pipe).  Bits 24..30 of S say "the last 4
//  bytes are the first 4 bytes (or all the bytes) of some pattern of bucket b"; bit 31 says "this
//  byte is NUL".
//
//  SLOW PATH (rare).  Lanes whose 32 start positions raised a flag append their group to a per-warp
//  shared-memory list.  When the list cannot take the next row's entries the warp drains it with
//  (nearly) every lane busy.
//    - simple entries (no packet boundary within reach, packet not yet NUL-terminated) carry their 36
//      bytes with them: the lane recomputes which start positions fired and walks the pattern trie
//      (a 16-bit copy in shared memory when it fits) from each of them, start-anchored, so a miss
//      dies after a byte or two;
//    - complex entries (a packet boundary inside the group or within pattern length of it) take the
//      general walk of the union automaton (the merged KMP DFAs, csrc/host/automaton.c), which
//      follows the offsets array and the reference's "text ends at the first NUL" rule (serial.c:191).
//  Either way every pattern occurrence that STARTS inside the group, lies inside one packet and has no
//  NUL before it in that packet is counted exactly once.  Counts go to shared-memory counters and
//  leave the block as one atomic per distinct pattern.
//
//  Packet boundaries and NULs are tracked per warp while streaming: a work item starts and ends on
//  packet boundaries, so "was there a NUL earlier in this packet" is known from the ballots of the
//  rows already scanned.  No separators, no padding and no second pass over the payload.
#include <algorithm>

#include "kmpb_device.cuh"

#ifndef KMPB_UN_THREADS
#define KMPB_UN_THREADS 640
#endif
#ifndef KMPB_UN_ITEM_KB
#define KMPB_UN_ITEM_KB 64
#endif
constexpr int UN_THREADS = KMPB_UN_THREADS; // one block per SM
constexpr int UN_WARPS = UN_THREADS / 32;
constexpr uint32_t UN_GRP = 32;                           // bytes per lane per row
constexpr uint32_t UN_ROW = 32 * UN_GRP;                  // bytes per warp row
constexpr uint32_t UN_ITEM_BYTES = KMPB_UN_ITEM_KB << 10; // target work-item size
constexpr uint32_t UN_QCAP = 32;                          // list entries per warp and kind
constexpr uint32_t UN_QS_WORDS = 12; // simple entry: 32 B group, 4 B lookahead, group index, NUL flag, pad (48 B)
constexpr uint32_t UN_QC_WORDS = 4;  // complex entry: group index, zone|dead, first/last packet of the item
constexpr uint32_t UN_LUT_BYTES = 256 * 256; // 256-byte row per byte value; lanes use the first 128 B
constexpr uint32_t UN_NOBOUND = 0xffffffffu;
constexpr uint32_t FULL = 0xffffffffu;

// Dynamic shared memory.  The LUT must start at a 64 KB-aligned shared address; the gap in front of it
// (63 KB when the dynamic window starts at 0x400, the usual case) holds the complex lists, per-warp
// scratch, the byte classes, the counters and the 16-bit trie; the simple lists follow the LUT.
constexpr uint32_t UN_QS_BYTES = UN_WARPS * UN_QCAP * UN_QS_WORDS * 4;
constexpr uint32_t UN_QC_BYTES = UN_WARPS * UN_QCAP * UN_QC_WORDS * 4;
constexpr uint32_t UN_SCRATCH_BYTES = UN_WARPS * 32;
constexpr uint32_t UN_FRONT_FIXED = UN_QC_BYTES + UN_SCRATCH_BYTES + 256 + 16;
constexpr uint32_t UN_FRONT_MAX = 60 * 1024; // what the gap is trusted to hold
constexpr size_t UN_SMEM_BYTES = 65536 + UN_LUT_BYTES + UN_QS_BYTES;

struct union_params {
    const uint8_t *bytes; // device pointer to absolute byte abs_base (abs_base % 512 == 0)
    uint64_t abs_base;
    const uint64_t *offsets; // [n_packets+1], absolute
    uint32_t n_packets;
    const uint32_t *items; // [n_items+1] first packet of each work item
    uint32_t n_items;
    uint32_t *work;         // [0] next item, [1] error flags
    const uint32_t *filter; // [256]
    const uint32_t *next;   // union DFA: next[state*n_class + class] = target | reports<<31
    const uint32_t *out_head;
    const uint32_t *out_id;
    const uint32_t *uniq_len;
    const uint32_t *trie;       // bare trie: child | pattern-ends-here<<31, 0 = no edge
    const uint32_t *state_term; // distinct pattern ending at a state
    const uint8_t *byte_class;
    uint32_t n_class, n_uniq, n_state, max_len;
    uint32_t counts_in_smem; // counters live in shared memory
    uint32_t vtab_in_smem;   // the hash verification tables live in shared memory
    const uint32_t *vtab;    // hash verification tables (automaton.c build_verify_tables)
    uint32_t vtab_words;
    uint32_t mul256; // the value 256, passed at run time so the shift-or-0xff compiles to an integer
                     // multiply-add on the FMA pipe instead of competing for the ALU pipe
    unsigned long long *uniq_counts;
};

// ---- work partition: item i = packets [items[i], items[i+1]) ---------------------------------
__global__ void kmpb_union_partition_kernel(const uint64_t *__restrict__ offsets, uint32_t n_packets,
                                            uint32_t n_items, uint64_t item_bytes, uint32_t *__restrict__ items,
                                            uint32_t *__restrict__ work)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { work[0] = 0; work[1] = 0; }
    if (i > n_items) return;
    if (i == n_items) { items[i] = n_packets; return; }
    // first packet whose start is >= offsets[0] + i * item_bytes
    const uint64_t target = offsets[0] + (uint64_t)i * item_bytes;
    uint32_t lo = 0, hi = n_packets;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (offsets[mid] < target) lo = mid + 1; else hi = mid;
    }
    items[i] = lo;
}

// ---- helpers -----------------------------------------------------------------------------------
struct grp { uint32_t w[8]; }; // one lane's 32 bytes of a row

// streaming 32-byte load: read once, keep it out of L1 (which holds the slow path's tables)
__device__ __forceinline__ void ld_stream32(const uint8_t *p, grp &g)
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(g.w[0]), "=r"(g.w[1]), "=r"(g.w[2]), "=r"(g.w[3]), "=r"(g.w[4]), "=r"(g.w[5]),
                   "=r"(g.w[6]), "=r"(g.w[7])
                 : "l"(p));
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
__device__ __forceinline__ uint32_t lds32v(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// 4 text bytes starting at byte `pos` of a queue entry (pos + 4 <= 36)
__device__ __forceinline__ uint32_t entry_window(uint32_t entry_sa, uint32_t pos)
{
    const uint32_t a = entry_sa + (pos & ~3u);
    return __funnelshift_r(lds32v(a), lds32v(a + 4), 8u * (pos & 3u));
}
__device__ __forceinline__ uint32_t saddr_of(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 0x80 in every byte of w that is zero (exact, no false positives above a zero byte)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t w)
{
    uint32_t t = (w & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | w | 0x7f7f7f7fu);
}
__device__ __forceinline__ uint32_t pack4(uint32_t z) { return (((z >> 7) * 0x00204081u) >> 21) & 0xfu; }
// bit i set when byte i of the 32-byte group is NUL
__device__ __forceinline__ uint32_t zero_mask32(const uint32_t *w)
{
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) m |= pack4(zero_bytes(w[i])) << (4 * i);
    return m;
}
// mask of lanes >= l (l may be 32)
__device__ __forceinline__ uint32_t lanes_ge(uint32_t l) { return l >= 32 ? 0u : ~((1u << l) - 1u); }

// block-wide shared tables of the slow path
struct slow_tables {
    uint32_t class_sa;  // shared address of the 256 byte classes
    uint32_t vtab_sa;   // shared address of the hash verification tables (vtab_in_smem)
    uint32_t *s_counts; // counters, or nullptr
};

__device__ __forceinline__ void count_hit(const union_params &p, const slow_tables &t, uint32_t u)
{
    if (t.s_counts) atomicAdd(&t.s_counts[u], 1u);
    else atomicAdd(p.uniq_counts + u, 1ull);
}

// filter word of byte `sel` of `word`: the PRMT builds the whole shared address
// [lutlane.b0 | byte | lutlane.b2 | lutlane.b3], lutlane = 64 KB-aligned LUT base + 4*lane
#define LUT_AT(word, sel) lds32(__byte_perm((word), lutlane, (sel)))
#define SEL0 0x7604
#define SEL1 0x7614
#define SEL2 0x7624
#define SEL3 0x7634
#define SA_NEXT(word, sel) (S = (S * mul + 255u) & LUT_AT(word, sel))
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
// same step, recording in bit `bit` of cm whether a candidate start fired
#define SV_STEP(word, sel, bit)                       \
    do {                                              \
        SA_NEXT(word, sel);                           \
        cm |= (S & 0x7f000000u) ? (1u << (bit)) : 0u; \
    } while (0)
#define SV_WORD(word, bit0)              \
    do {                                 \
        SV_STEP(word, SEL0, (bit0));     \
        SV_STEP(word, SEL1, (bit0) + 1); \
        SV_STEP(word, SEL2, (bit0) + 2); \
        SV_STEP(word, SEL3, (bit0) + 3); \
    } while (0)

// Slow path, simple entry: the group's 32 bytes + 4 bytes of lookahead sit in shared memory at
// `entry`; no packet boundary lies within reach of a match starting in the group and no NUL precedes
// the group in its packet.  Start-anchored trie walk from every start position that fired.
__device__ __noinline__ void verify_simple(const union_params &p, const slow_tables &t, const uint32_t *entry,
                                           uint32_t lutlane, uint32_t mul)
{
    uint32_t w[8];
    *reinterpret_cast<uint4 *>(w) = *reinterpret_cast<const uint4 *>(entry);
    *reinterpret_cast<uint4 *>(w + 4) = *reinterpret_cast<const uint4 *>(entry + 4);
    const uint32_t la = entry[8], g32 = entry[9], has_nul = entry[10];
    uint32_t S, cm = 0;
    S = LUT_AT(w[0], SEL0) & 0x808080ffu;
    SA_NEXT(w[0], SEL1);
    SA_NEXT(w[0], SEL2);
    SV_STEP(w[0], SEL3, 0);
    SV_WORD(w[1], 1); SV_WORD(w[2], 5); SV_WORD(w[3], 9); SV_WORD(w[4], 13);
    SV_WORD(w[5], 17); SV_WORD(w[6], 21); SV_WORD(w[7], 25);
    SV_STEP(la, SEL0, 29); SV_STEP(la, SEL1, 30); SV_STEP(la, SEL2, 31);
    if (has_nul) { // starts at or after the group's first NUL are dead (serial.c:191)
        const uint32_t zm = zero_mask32(w);
        if (zm) cm &= (1u << (__ffs(zm) - 1)) - 1u;
    }
    const uint32_t entry_sa = saddr_of(entry);
    const uint32_t ncls = p.n_class;
    if (p.vtab_in_smem) {
        // Hash verification (tables of csrc/host/automaton.c build_verify_tables, copied to shared
        // memory): the first min(len,4) bytes of every pattern are a key in the table of that key
        // length; a hit is confirmed by comparing the remaining pattern words.  No per-byte walk: two or
        // three dependent shared-memory reads per candidate instead of one per pattern byte.
        const uint32_t vt = t.vtab_sa;
        const uint32_t lens = lds32(vt + 44), rec_sa = vt + 4u * lds32(vt + 36), pat_sa = vt + 4u * lds32(vt + 40);
        const uint8_t *gb = p.bytes + 32ull * g32;
        while (cm) {
            const uint32_t i = __ffs(cm) - 1;
            cm &= cm - 1;
            const uint32_t x0 = entry_window(entry_sa, i);
#pragma unroll
            for (uint32_t L = 1; L <= 4; L++) {
                if (!((lens >> (L - 1)) & 1u)) continue;
                const uint32_t key = L == 4 ? x0 : x0 & ((1u << (8 * L)) - 1u);
                const uint32_t mask = lds32(vt + 16u + 4u * L), tab_sa = vt + 4u * lds32(vt + 4u * L);
                for (uint32_t slot = ((key * 0x9e3779b1u) >> 12) & mask;; slot = (slot + 1) & mask) {
                    const uint2 e = lds64(tab_sa + 8u * slot);
                    if (e.y == 0xffffffffu) break;
                    if (e.x != key) continue;
                    for (uint32_t u = e.y; u != 0xffffffffu; u = lds32(rec_sa + 12u * u + 8u)) {
                        const uint32_t m = lds32(rec_sa + 12u * u), pw_sa = pat_sa + 4u * lds32(rec_sa + 12u * u + 4u);
                        bool same = true;
                        for (uint32_t j = 4; j < m && same; j += 4) { // pattern bytes j..j+3 against text bytes i+j..
                            const uint32_t pw = lds32(pw_sa + j), rem = m - j;
                            if (i + j + 4 <= 36) {
                                const uint32_t diff = entry_window(entry_sa, i + j) ^ pw;
                                same = (rem >= 4 ? diff : diff & ((1u << (8 * rem)) - 1u)) == 0;
                            } else { // past the bytes carried along: byte by byte, entry first, then global memory
                                for (uint32_t b = 0; b < 4 && b < rem && same; b++) {
                                    const uint32_t pos = i + j + b;
                                    const uint32_t c = pos < 36 ? lds8v(entry_sa + pos) : (uint32_t)gb[pos];
                                    same = c == ((pw >> (8 * b)) & 0xffu);
                                }
                            }
                        }
                        if (same) count_hit(p, t, u);
                    }
                    break;
                }
            }
        }
    } else {
        const uint8_t *gb = p.bytes + 32ull * g32;
        const uint32_t *trie = p.trie;
        const uint32_t *term = p.state_term;
        while (cm) {
            const uint32_t i = __ffs(cm) - 1;
            cm &= cm - 1;
            uint32_t node = 0;
            for (uint32_t k = i;; k++) {
                const uint32_t c = k < 36 ? lds8v(entry_sa + k) : (uint32_t)gb[k];
                const uint32_t e = __ldg(trie + node * ncls + lds8v(t.class_sa + c));
                if (e == 0) break;
                node = e & 0x7fffffffu;
                if (e >> 31) count_hit(p, t, __ldg(term + node));
            }
        }
    }
}

// Slow path, complex entry: general walk.  g32 = group index (32-byte units from abs_base); the
// item holds packets [ks, ke); zones: 0 = before the item's first packet (dead), j = packet ks+j-1,
// > ke-ks = after the item's last packet (dead).  zone = the zone holding the group's first byte,
// dead = a NUL precedes it inside that packet.  Counts every pattern occurrence that starts in the
// group, lies inside one packet of the item and has no NUL before it in that packet.
__device__ __noinline__ void verify_complex(const union_params &p, const slow_tables &t, uint32_t g32, uint32_t zone,
                                            bool dead, uint32_t ks, uint32_t ke)
{
    const uint64_t *off = p.offsets + ks;
    const uint32_t nbound = ke - ks;
    const uint64_t g = p.abs_base + (uint64_t)UN_GRP * g32;
    uint64_t limit = g + (UN_GRP - 1) + p.max_len; // one past the last byte a match starting at g+31 can touch
    if (limit > off[nbound]) limit = off[nbound];
    uint64_t nb = zone <= nbound ? off[zone] : ~0ull;
    const uint8_t *text = p.bytes - p.abs_base;
    uint32_t state = 0;
    for (uint64_t pos = g; pos < limit; pos++) {
        while (pos == nb) { // crossing into the next packet (or out of the item)
            state = 0;
            dead = false;
            zone++;
            nb = zone <= nbound ? off[zone] : ~0ull;
        }
        if (zone > nbound) break;
        const uint32_t c = text[pos];
        if (c == 0) dead = true; // strlen() in kmp_matcher stops here for the rest of the packet
        if (dead) {
            if (nb >= limit) break; // nothing can revive before the walk ends
            state = 0;
            continue;
        }
        const uint32_t e = __ldg(p.next + state * p.n_class + lds8v(t.class_sa + c));
        state = e & 0x7fffffffu;
        if (e >> 31) {
            const uint32_t o1 = __ldg(p.out_head + state + 1);
            for (uint32_t o = __ldg(p.out_head + state); o < o1; o++) {
                const uint32_t u = __ldg(p.out_id + o);
                if (pos + 1 - __ldg(p.uniq_len + u) < g + UN_GRP) // start >= g holds: the walk began at g in the root
                    count_hit(p, t, u);
            }
        }
        if (pos >= g + (UN_GRP - 1) && state == 0) break; // no match in flight that started inside the group
    }
}

// per-warp streaming state
struct warp_state {
    // item
    const uint8_t *text; // byte 0 of the item's first row
    const uint64_t *off; // item boundary j is off[j] - row0
    uint64_t row0;
    uint32_t nbound, e_rel, load_end, g32_0, ks, ke;
    // zone tracking (warp-uniform): zone = boundaries crossed so far; 0 = before the first packet
    uint32_t zone, nb, nb_next;
    bool dead;
    // boundary window: the item's boundaries (relative to row0) are fetched 32 at a time, one per lane;
    // bcur holds boundaries [bbase, bbase+32), bnxt the 32 after them (already in flight)
    uint32_t bcur, bnxt, bbase;
    // pending entries
    uint32_t qs_n, qc_n;
};

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
    const uint32_t front_bytes = UN_FRONT_FIXED + counts_bytes + vtab_bytes;
    if (lut_off < front_bytes || lut_off + UN_LUT_BYTES + UN_QS_BYTES > dyn_size) {
        // unexpected shared-memory window base: refuse rather than compute something wrong
        if (threadIdx.x == 0) atomicOr(&p.work[1], 2u);
        return;
    }
    uint8_t *lut = smem + lut_off;
    uint32_t *qs_all = reinterpret_cast<uint32_t *>(lut + UN_LUT_BYTES);
    uint32_t *qc_all = reinterpret_cast<uint32_t *>(smem);
    uint8_t *scratch_all = smem + UN_QC_BYTES;
    uint8_t *s_class = scratch_all + UN_SCRATCH_BYTES;
    uint32_t *s_lut_saddr = reinterpret_cast<uint32_t *>(s_class + 256);
    uint32_t *s_counts = reinterpret_cast<uint32_t *>(s_lut_saddr + 4);
    uint32_t *s_vtab = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(s_counts) + counts_bytes);

    for (uint32_t i = threadIdx.x; i < 256 * 32; i += UN_THREADS)
        reinterpret_cast<uint32_t *>(lut)[(i >> 5) * 64 + (i & 31)] = p.filter[i >> 5];
    for (uint32_t i = threadIdx.x; i < 256; i += UN_THREADS) s_class[i] = p.byte_class[i];
    if (p.counts_in_smem)
        for (uint32_t i = threadIdx.x; i < p.n_uniq; i += UN_THREADS) s_counts[i] = 0;
    if (p.vtab_in_smem)
        for (uint32_t i = threadIdx.x; i < p.vtab_words; i += UN_THREADS) s_vtab[i] = p.vtab[i];
    if (threadIdx.x == 0) *s_lut_saddr = dyn_saddr + lut_off;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    // read back through shared memory so that no LUT load can be scheduled above the barrier
    const uint32_t lutlane = *s_lut_saddr + (lane << 2);
    const uint32_t mul = p.mul256;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t *qs = qs_all + warp * (UN_QCAP * UN_QS_WORDS);
    uint32_t *qc = qc_all + warp * (UN_QCAP * UN_QC_WORDS);
    uint32_t *scratch = reinterpret_cast<uint32_t *>(scratch_all + warp * 32);
    const uint32_t scratch_sa = saddr_of(scratch);
    slow_tables st;
    st.class_sa = saddr_of(s_class);
    st.vtab_sa = saddr_of(s_vtab);
    st.s_counts = p.counts_in_smem ? s_counts : nullptr;
    const uint32_t reach = (UN_GRP - 1) + p.max_len;
    warp_state w;
    w.qs_n = w.qc_n = 0;

    auto drain_simple = [&]() {
        __syncwarp();
        if (lane < w.qs_n) verify_simple(p, st, qs + lane * UN_QS_WORDS, lutlane, mul);
        w.qs_n = 0;
        __syncwarp();
    };
    auto drain_complex = [&]() {
        __syncwarp();
        if (lane < w.qc_n) {
            const uint4 e = *reinterpret_cast<const uint4 *>(qc + lane * UN_QC_WORDS);
            verify_complex(p, st, e.x, e.y & 0x7fffffffu, (e.y >> 31) != 0, e.z, e.w);
        }
        w.qc_n = 0;
        __syncwarp();
    };

    // this lane's share of the 32 boundaries starting at boundary index `first`
    auto boundary_load = [&](uint32_t first) -> uint32_t {
        const uint32_t j = first + lane;
        return j <= w.nbound ? (uint32_t)(w.off[j] - w.row0) : UN_NOBOUND;
    };
    // after w.zone++: the boundary after the one that now ends the current zone (index zone+1)
    auto boundary_after_cross = [&]() -> uint32_t {
        if (w.zone >= w.bbase + 32) {
            w.bcur = w.bnxt;
            w.bbase += 32;
            w.bnxt = boundary_load(w.bbase + 32);
        }
        const uint32_t r = w.zone + 1 - w.bbase; // 1..32
        return __shfl_sync(FULL, r < 32 ? w.bcur : w.bnxt, r & 31);
    };

    // one 1024-byte row: `cur` is scanned, `nxt` supplies lane 31's lookahead
    auto scan_row = [&](const grp &cur, const grp &nxt, const uint32_t row, const uint32_t g) {
        // 3 bytes of lookahead: first word of the next group (next lane, or lane 0 of the next row)
        const uint32_t la = __shfl_sync(FULL, lane == 0 ? nxt.w[0] : cur.w[0], (lane + 1) & 31);

        // ---- shift-and filter over 35 bytes ------------------------------------------------------
        uint32_t accA, accB, accC;
#ifdef KMPB_UN_TWO_CHAINS
        // two independent 19-byte chains (bytes 0..18 and 16..34): the dependent IMAD -> LOP3 steps of
        // one fill the latency gaps of the other, at the price of 3 extra steps
        {
            uint32_t S, T, tA, tB, tC;
            S = LUT_AT(cur.w[0], SEL0) & 0x808080ffu;
            T = LUT_AT(cur.w[4], SEL0) & 0x808080ffu;
            accA = S; tA = T;
#define SB_STEP(word, sel, acc) do { T = (T * mul + 255u) & LUT_AT(word, sel); acc |= T; } while (0)
            SA_STEP(cur.w[0], SEL1, accA); SB_STEP(cur.w[4], SEL1, tA);
            SA_STEP(cur.w[0], SEL2, accA); SB_STEP(cur.w[4], SEL2, tA);
            accB = 0; tB = 0;
            SA_STEP(cur.w[0], SEL3, accB); SB_STEP(cur.w[4], SEL3, tB);
            SA_STEP(cur.w[1], SEL0, accB); SB_STEP(cur.w[5], SEL0, tB);
            SA_STEP(cur.w[1], SEL1, accB); SB_STEP(cur.w[5], SEL1, tB);
            SA_STEP(cur.w[1], SEL2, accB); SB_STEP(cur.w[5], SEL2, tB);
            SA_STEP(cur.w[1], SEL3, accB); SB_STEP(cur.w[5], SEL3, tB);
            SA_STEP(cur.w[2], SEL0, accB); SB_STEP(cur.w[6], SEL0, tB);
            SA_STEP(cur.w[2], SEL1, accB); SB_STEP(cur.w[6], SEL1, tB);
            SA_STEP(cur.w[2], SEL2, accB); SB_STEP(cur.w[6], SEL2, tB);
            SA_STEP(cur.w[2], SEL3, accB); SB_STEP(cur.w[6], SEL3, tB);
            SA_STEP(cur.w[3], SEL0, accB); SB_STEP(cur.w[7], SEL0, tB);
            SA_STEP(cur.w[3], SEL1, accB); SB_STEP(cur.w[7], SEL1, tB);
            SA_STEP(cur.w[3], SEL2, accB); SB_STEP(cur.w[7], SEL2, tB);
            SA_STEP(cur.w[3], SEL3, accB); SB_STEP(cur.w[7], SEL3, tB);
            accC = 0; tC = 0;
            SA_STEP(cur.w[4], SEL0, accC); SB_STEP(la, SEL0, tC);
            SA_STEP(cur.w[4], SEL1, accC); SB_STEP(la, SEL1, tC);
            SA_STEP(cur.w[4], SEL2, accC); SB_STEP(la, SEL2, tC);
#undef SB_STEP
            accA |= tA; accB |= tB; accC |= tC;
        }
#else
        uint32_t S;
        S = LUT_AT(cur.w[0], SEL0) & 0x808080ffu; // no history: only the NUL stage is pre-armed
        accA = S;
        SA_STEP(cur.w[0], SEL1, accA);
        SA_STEP(cur.w[0], SEL2, accA);
        accB = 0;
        SA_STEP(cur.w[0], SEL3, accB);
        SA_WORD(cur.w[1], accB);
        SA_WORD(cur.w[2], accB);
        SA_WORD(cur.w[3], accB);
        SA_WORD(cur.w[4], accB);
        SA_WORD(cur.w[5], accB);
        SA_WORD(cur.w[6], accB);
        SA_WORD(cur.w[7], accB);
        accC = 0;
        SA_STEP(la, SEL0, accC);
        SA_STEP(la, SEL1, accC);
        SA_STEP(la, SEL2, accC);
#endif
        const bool nul = ((accA | accB) >> 31) != 0;          // a NUL among my 32 bytes
        const bool cand = ((accB | accC) & 0x7f000000u) != 0; // a candidate start among my 32 positions
        const uint32_t nulm = __ballot_sync(FULL, nul);
        const uint32_t candm = __ballot_sync(FULL, cand);

        // ---- which lanes push, and into which list ------------------------------------------------
        const uint32_t row_end = row + UN_ROW;
        uint32_t ms, mc;      // lanes pushing a simple / complex entry
        uint32_t kz = w.zone; // zone of my group's first byte
        bool d0 = false;      // my packet already saw a NUL before my group
#ifdef KMPB_ABLATE_BOUNDARY // measurement only (wrong counts): cost of the boundary path
        if (w.nb >= row_end || w.nb != 0x12345678u) {
#else
        if (w.nb >= row_end) { // no packet boundary inside this row: everything is warp-uniform
#endif
            const uint32_t low = nulm & (0u - nulm);
            const uint32_t deadm = w.dead ? FULL : (nulm ? ~((low << 1) - 1u) : 0u); // lanes above the first NUL lane
            const uint32_t alive = candm & ~deadm;
            // lanes whose reach (group start + 31 + longest pattern) crosses the next boundary
            const uint32_t nearm = w.nb == UN_NOBOUND ? 0u : lanes_ge(((w.nb - reach - row) >> 5) + 1u);
            ms = alive & ~nearm;
            mc = alive & nearm;
            w.dead = w.dead || nulm != 0;
        } else if (w.nb_next >= row_end && w.zone + 1 <= w.nbound) {
            // exactly one packet boundary inside this row and a packet follows it (the usual case for
            // packets longer than a row): warp-uniform mask arithmetic, no loop
            const uint32_t r = w.nb - row, lb = r >> 5, ob = r & (UN_GRP - 1);
            const uint32_t lbit = 1u << lb, below = lbit - 1u;
            // old packet: lanes below lb
            const uint32_t nul_old = nulm & below;
            const uint32_t low_old = nul_old & (0u - nul_old);
            const uint32_t dead_old = w.dead ? FULL : (nul_old ? ~((low_old << 1) - 1u) : 0u);
            const uint32_t alive_old = candm & below & ~dead_old;
            const uint32_t near_old = lanes_ge(r >= reach ? ((r - reach) >> 5) + 1u : 0u);
            ms = alive_old & ~near_old;
            mc = alive_old & near_old;
            // new packet: lanes above lb, and lane lb itself when the boundary is exactly at its start
            const uint32_t newm = ob ? ~(below | lbit) : ~below;
            bool nafter = false; // the new packet has a NUL inside lane lb's group
            if (ob) {
                mc |= candm & lbit; // the group holding the boundary: the general walk sorts it out
                if (lane == lb) d0 = (dead_old & lbit) != 0;
                if (nulm & lbit) {
                    // is one of lane lb's NULs at or after the boundary?  lane lb spreads its 32 bytes
                    // through shared memory, lane j looks at byte j
                    __syncwarp();
                    if (lane == lb) {
                        *reinterpret_cast<uint4 *>(scratch) = make_uint4(cur.w[0], cur.w[1], cur.w[2], cur.w[3]);
                        *reinterpret_cast<uint4 *>(scratch + 4) = make_uint4(cur.w[4], cur.w[5], cur.w[6], cur.w[7]);
                    }
                    __syncwarp();
                    nafter = __ballot_sync(FULL, lds8v(scratch_sa + lane) == 0 && lane >= ob) != 0;
                }
            }
            const uint32_t nul_new = nulm & newm;
            const uint32_t low_new = nul_new & (0u - nul_new);
            const uint32_t dead_new = nafter ? FULL : (nul_new ? ~((low_new << 1) - 1u) : 0u);
            const uint32_t alive_new = candm & newm & ~dead_new;
            const uint32_t near_new = w.nb_next == UN_NOBOUND ? 0u : lanes_ge(((w.nb_next - reach - row) >> 5) + 1u);
            ms |= alive_new & ~near_new;
            mc |= alive_new & near_new;
            if ((newm >> lane) & 1u) kz++;
            w.dead = nafter || nul_new != 0;
            w.zone++;
            w.nb = w.nb_next;
            w.nb_next = boundary_after_cross();
        } else {
            uint32_t bin = 0, bat = 0; // lanes with a boundary inside their group / exactly at its start
            uint32_t endm = 0;         // lanes past the item's last packet
            uint32_t nafter = 0;       // lanes whose group has a NUL after its last inner boundary
            uint32_t my_nb = UN_NOBOUND;
            bool ended = false;
            while (w.nb < row_end) {
                const uint32_t lb = (w.nb - row) >> 5, ob = (w.nb - row) & (UN_GRP - 1);
                if (w.nb <= g) kz++;
                else if (my_nb == UN_NOBOUND) my_nb = w.nb;
                if (ob) {
                    bin |= 1u << lb;
                    if ((nulm >> lb) & 1u) {
                        // is one of lane lb's NULs at or after the boundary?  lane lb spreads its 32
                        // bytes through shared memory, lane j looks at byte j
                        __syncwarp();
                        if (lane == lb) {
                            *reinterpret_cast<uint4 *>(scratch) = make_uint4(cur.w[0], cur.w[1], cur.w[2], cur.w[3]);
                            *reinterpret_cast<uint4 *>(scratch + 4) = make_uint4(cur.w[4], cur.w[5], cur.w[6], cur.w[7]);
                        }
                        __syncwarp();
                        const bool z = lds8v(scratch_sa + lane) == 0 && lane >= ob;
                        if (__ballot_sync(FULL, z)) nafter |= 1u << lb;
                        else nafter &= ~(1u << lb);
                    } else {
                        nafter &= ~(1u << lb);
                    }
                } else {
                    bat |= 1u << lb;
                }
                w.zone++;
                if (w.zone > w.nbound) { // past the item's last packet
                    endm = ob ? ~((2u << lb) - 1u) : ~((1u << lb) - 1u);
                    ended = true;
                    w.nb = UN_NOBOUND;
                    break;
                }
                w.nb = w.nb_next;
                w.nb_next = boundary_after_cross();
            }
            if (my_nb == UN_NOBOUND) my_nb = w.nb;
            const uint32_t before = (bin & lt) | (bat & (lt | (1u << lane))); // boundaries at or before my first byte
            if (before == 0) {
                d0 = w.dead || (nulm & lt) != 0;
            } else {
                const uint32_t j = 31u - __clz(before);
                if (j == lane) d0 = false; // my group starts a packet
                else if ((bin >> j) & 1u) d0 = ((nafter >> j) & 1u) != 0 || (nulm & lt & ~((2u << j) - 1u)) != 0;
                else d0 = (nulm & lt & ~((1u << j) - 1u)) != 0;
            }
            if ((endm >> lane) & 1u) d0 = true;
            const bool push_s = cand && !d0 && my_nb >= g + reach;
            const bool push_c = cand && !push_s && (!d0 || my_nb < g + UN_GRP); // a boundary inside the group can revive it
            ms = __ballot_sync(FULL, push_s);
            mc = __ballot_sync(FULL, push_c);
            // state of the zone the row ends in
            const uint32_t all = bin | bat;
            const uint32_t j = 31u - __clz(all); // all != 0: at least one boundary was crossed
            if (ended) w.dead = true;
            else if ((bin >> j) & 1u) w.dead = ((nafter >> j) & 1u) != 0 || (nulm & ~((2u << j) - 1u)) != 0;
            else w.dead = (nulm & ~((1u << j) - 1u)) != 0;
        }

        // ---- append flagged groups; drain first when the list cannot take them ----------------------
#ifdef KMPB_ABLATE_SLOW_PATH // measurement only (wrong counts): how fast is the fast path alone?
        if (ms == 0x12345678u && mc == 0x9abcdef0u)
#endif
        if (ms) {
            const uint32_t n = __popc(ms);
            if (w.qs_n + n > UN_QCAP) drain_simple();
            if ((ms >> lane) & 1u) {
                uint32_t *e = qs + (w.qs_n + __popc(ms & lt)) * UN_QS_WORDS;
                *reinterpret_cast<uint4 *>(e) = make_uint4(cur.w[0], cur.w[1], cur.w[2], cur.w[3]);
                *reinterpret_cast<uint4 *>(e + 4) = make_uint4(cur.w[4], cur.w[5], cur.w[6], cur.w[7]);
                *reinterpret_cast<uint4 *>(e + 8) = make_uint4(la, w.g32_0 + (g >> 5), nul ? 1u : 0u, 0u);
            }
            w.qs_n += n;
        }
        if (mc) {
            const uint32_t n = __popc(mc);
            if (w.qc_n + n > UN_QCAP) drain_complex();
            if ((mc >> lane) & 1u) {
                uint32_t *e = qc + (w.qc_n + __popc(mc & lt)) * UN_QC_WORDS;
                *reinterpret_cast<uint4 *>(e) = make_uint4(w.g32_0 + (g >> 5), kz | (d0 ? 0x80000000u : 0u), w.ks, w.ke);
            }
            w.qc_n += n;
        }
    };

    auto load_row = [&](grp &v, uint32_t g) {
        if (g < w.load_end) {
            ld_stream32(w.text + g, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v.w[i] = 0;
        }
    };

    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&p.work[0], 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= p.n_items) break;
        w.ks = p.items[item];
        w.ke = p.items[item + 1];
        if (w.ks >= w.ke) continue;
        const uint64_t b_abs = p.offsets[w.ks], e_abs = p.offsets[w.ke];
        if (b_abs == e_abs) continue;
        if (e_abs - b_abs >= (1ull << 31)) { // a packet over 2 GiB: outside the documented limits
            if (lane == 0) atomicOr(&p.work[1], 1u);
            continue;
        }
        w.row0 = b_abs & ~127ull; // absolute position of the item's first row
        w.text = p.bytes + (w.row0 - p.abs_base);
        w.off = p.offsets + w.ks;
        w.nbound = w.ke - w.ks;
        w.e_rel = (uint32_t)(e_abs - w.row0);
        w.load_end = (w.e_rel + (UN_GRP - 1)) & ~(UN_GRP - 1);
        w.g32_0 = (uint32_t)((w.row0 - p.abs_base) >> 5);
        w.zone = 0;
        w.bbase = 0;
        w.bcur = boundary_load(0);
        w.bnxt = boundary_load(32);
        w.nb = (uint32_t)(b_abs - w.row0);
        w.nb_next = __shfl_sync(FULL, w.bcur, 1);
        w.dead = true;

        // three row buffers (scanned / lookahead / in flight): a row is requested two row-times before
        // it is scanned, one before it serves as lookahead.  The buffers rotate by register moves (16
        // IMAD.MOVs on the otherwise idle FMA pipe) rather than by unrolling the row body three times:
        // the unrolled kernel no longer fits the instruction cache (ncu: stall_no_instruction 3.6/issue).
        uint32_t g = lane * UN_GRP;
        grp a, b, c;
        load_row(a, g);
        load_row(b, g + UN_ROW);
#pragma unroll 1
        for (uint32_t row = 0; row < w.e_rel; row += UN_ROW, g += UN_ROW) {
            load_row(c, g + 2 * UN_ROW);
            scan_row(a, b, row, g);
            a = b;
            b = c;
        }
    }
    // leftovers
    drain_simple();
    drain_complex();

    __syncthreads();
    if (p.counts_in_smem)
        for (uint32_t u = threadIdx.x; u < p.n_uniq; u += UN_THREADS)
            if (s_counts[u]) atomicAdd(p.uniq_counts + u, (unsigned long long)s_counts[u]);
}

// ---- host side ------------------------------------------------------------------------------------

// scratch for batches of up to max_batch_bytes: work counters and the item table, one set per slot
int kmpb_union_scratch(kmpb_ctx *ctx, uint64_t max_batch_bytes)
{
    size_t need = (size_t)(max_batch_bytes / UN_ITEM_BYTES) + 3;
    if (need <= ctx->items_cap && ctx->d_items && ctx->d_work) return KMPB_OK;
    cudaFree(ctx->d_items);
    ctx->d_items = nullptr;
    ctx->items_cap = 0;
    if (!ctx->d_work) KMPB_CUDA(cudaMalloc((void **)&ctx->d_work, KMPB_COPY_STREAMS * 4 * sizeof(uint32_t)));
    KMPB_CUDA(cudaMalloc((void **)&ctx->d_items, (size_t)KMPB_COPY_STREAMS * need * sizeof(uint32_t)));
    ctx->items_cap = need;
    return KMPB_OK;
}

int kmpb_launch_union(kmpb_ctx *ctx, const kmpb_batch &b, int slot, uint64_t *d_uniq_counts, cudaStream_t stream)
{
    const kmpb_tables &h = ctx->host;
    if (h.n_uniq == 0 || b.n_packets == 0 || b.end_byte == b.first_byte) return KMPB_OK;
    if (b.n_packets >= (1ull << 31)) return kmpb_fail(KMPB_ELIMIT, "more than 2^31-1 packets in one batch");
    if ((b.abs_base & 511) || ((uintptr_t)b.d_bytes & 31))
        return kmpb_fail(KMPB_EINVAL, "payload buffer must be 32-byte aligned");
    if (b.end_byte - b.abs_base >= (1ull << 36))
        return kmpb_fail(KMPB_ELIMIT, "more than 64 GiB of payload in one batch");
    const uint64_t span = b.end_byte - b.first_byte;
    const uint32_t n_items = (uint32_t)((span + UN_ITEM_BYTES - 1) / UN_ITEM_BYTES);
    if ((size_t)n_items + 1 > ctx->items_cap) return kmpb_fail(KMPB_ESTATE, "union scratch too small");
    uint32_t *d_items = ctx->d_items + (size_t)slot * ctx->items_cap;
    uint32_t *d_work = ctx->d_work + slot * 4;

    // what fits in the shared-memory gap in front of the LUT: counters first, then the 16-bit trie
    uint32_t front = UN_FRONT_FIXED;
    const bool counts_in_smem = front + 4ull * h.n_uniq + 16 <= UN_FRONT_MAX;
    if (counts_in_smem) front += (4u * h.n_uniq + 15u) & ~15u;
    const bool vtab_in_smem = front + 4ull * h.vtab_words <= UN_FRONT_MAX;
    if (!ctx->attr_union_set) {
        KMPB_CUDA(cudaFuncSetAttribute(kmpb_union_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UN_SMEM_BYTES));
        ctx->attr_union_set = true;
    }
    kmpb_union_partition_kernel<<<(n_items + 1 + 255) / 256, 256, 0, stream>>>(b.d_offsets, (uint32_t)b.n_packets, n_items,
                                                                              UN_ITEM_BYTES, d_items, d_work);
    union_params p;
    p.bytes = b.d_bytes;
    p.abs_base = b.abs_base;
    p.offsets = b.d_offsets;
    p.n_packets = (uint32_t)b.n_packets;
    p.items = d_items;
    p.n_items = n_items;
    p.work = d_work;
    p.filter = ctx->dev.filter;
    p.next = ctx->dev.next;
    p.out_head = ctx->dev.out_head;
    p.out_id = ctx->dev.out_id;
    p.uniq_len = ctx->dev.uniq_len;
    p.trie = ctx->dev.trie;
    p.state_term = ctx->dev.state_term;
    p.byte_class = ctx->dev.byte_class;
    p.n_class = h.n_class;
    p.n_uniq = h.n_uniq;
    p.n_state = h.n_state;
    p.max_len = h.max_len;
    p.counts_in_smem = counts_in_smem ? 1u : 0u;
    p.vtab_in_smem = vtab_in_smem ? 1u : 0u;
    p.vtab = ctx->dev.vtab;
    p.vtab_words = h.vtab_words;
    p.mul256 = 256u;
    p.uniq_counts = (unsigned long long *)d_uniq_counts;
    const uint32_t warps_needed = n_items;
    int grid = (int)std::min<uint32_t>((uint32_t)ctx->sm_count, (warps_needed + UN_WARPS - 1) / UN_WARPS);
    if (grid < 1) grid = 1;
    if (ctx->profile) KMPB_CUDA(cudaEventRecord(ctx->ev_kernel[0], stream));
    kmpb_union_kernel<<<grid, UN_THREADS, UN_SMEM_BYTES, stream>>>(p);
    if (ctx->profile) KMPB_CUDA(cudaEventRecord(ctx->ev_kernel[1], stream));
    ctx->launches += 2;
    KMPB_CUDA(cudaGetLastError());
    return KMPB_OK;
}
