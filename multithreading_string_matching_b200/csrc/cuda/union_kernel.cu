// union_kernel.cu -- the union engine (KMPB_ENGINE_UNION): every payload byte is read from HBM once.
//
// Two levels.
//
//  FAST PATH (every byte).  A warp streams work items -- runs of whole packets, ~32 KB of the flat
//  CSR byte buffer -- in rows of 512 contiguous bytes: one coalesced 16-byte load per lane, two rows
//  in flight ahead of the one being scanned.  Each lane pushes its 16 bytes (+3 bytes of lookahead
//  from its neighbour, by shuffle) through a 4-byte-deep shift-and filter over 8 buckets:
//      S = ((S << 8) | 0xff) & filter[byte]
//  filter[] lives in shared memory in a bank-private layout (byte address = byte*256 + lane*4), so the
//  one lookup per byte never bank-conflicts and its address is a single PRMT.  Bits 24..30 of S say
//  "the last 4 bytes are the first 4 bytes (or all the bytes) of some pattern of bucket b"; bit 31
//  says "this byte is NUL".
//
//  SLOW PATH (rare).  Lanes whose 16 start positions raised a flag push their group into a per-warp
//  shared-memory ring.  When 32 entries have gathered the warp drains them with every lane busy.
//    - simple entries (no packet boundary within reach, packet not yet NUL-terminated) carry their 20
//      bytes with them: the lane recomputes which start positions fired and walks the pattern trie
//      from each of them (start-anchored, so a miss dies after a byte or two);
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

constexpr int UN_THREADS = 1024; // one block per SM
constexpr int UN_WARPS = UN_THREADS / 32;
constexpr uint32_t UN_ROW = 512;             // bytes per warp row
constexpr uint32_t UN_ITEM_BYTES = 32 << 10; // target work-item size
constexpr uint32_t UN_QCAP = 64;             // ring entries per warp and kind
constexpr uint32_t UN_QS_WORDS = 8;          // simple entry: 16 B group, 4 B lookahead, group index, pad
constexpr uint32_t UN_QC_WORDS = 4;          // complex entry: group index, zone|dead, first/last packet of the item
constexpr uint32_t UN_LUT_BYTES = 256 * 256; // 256-byte row per byte value; lanes use the first 128 B
constexpr uint32_t UN_NOBOUND = 0xffffffffu;
constexpr uint32_t FULL = 0xffffffffu;

constexpr size_t UN_SMEM_FIXED = UN_LUT_BYTES + 256 + (size_t)UN_WARPS * UN_QCAP * (UN_QS_WORDS + UN_QC_WORDS) * 4;

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
    uint32_t n_class, n_uniq, max_len;
    uint32_t counts_in_smem;
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
// 0x80 in every byte of w that is zero (exact, no false positives above a zero byte)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t w)
{
    uint32_t t = (w & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | w | 0x7f7f7f7fu);
}
__device__ __forceinline__ uint32_t pack4(uint32_t z) { return (((z >> 7) * 0x00204081u) >> 21) & 0xfu; }
// bit i set when byte i of the 16-byte group is NUL
__device__ __forceinline__ uint32_t zero_mask16(const uint4 &v)
{
    return pack4(zero_bytes(v.x)) | pack4(zero_bytes(v.y)) << 4 | pack4(zero_bytes(v.z)) << 8 |
           pack4(zero_bytes(v.w)) << 12;
}

__device__ __forceinline__ void count_hit(const union_params &p, uint32_t *s_counts, uint32_t u)
{
    if (s_counts) atomicAdd(&s_counts[u], 1u);
    else atomicAdd(p.uniq_counts + u, 1ull);
}

#define LUT_AT(word, sel) (*reinterpret_cast<const uint32_t *>(lut + __byte_perm((word), laneoff, (sel))))
#define SA_STEP(word, sel, acc)                \
    do {                                       \
        const uint32_t m_ = LUT_AT(word, sel); \
        S = ((S << 8) | 0xffu) & m_;           \
        acc |= S;                              \
    } while (0)
#define SA_WORD(word, acc)          \
    do {                            \
        SA_STEP(word, 0x5504, acc); \
        SA_STEP(word, 0x5514, acc); \
        SA_STEP(word, 0x5524, acc); \
        SA_STEP(word, 0x5534, acc); \
    } while (0)
// same step, recording in bit `bit` of cm whether a candidate start fired
#define SV_STEP(word, sel, bit)                              \
    do {                                                     \
        const uint32_t m_ = LUT_AT(word, sel);               \
        S = ((S << 8) | 0xffu) & m_;                         \
        cm |= (S & 0x7f000000u) ? (1u << (bit)) : 0u;        \
    } while (0)

// Slow path, simple entry: the group's 16 bytes + 4 bytes of lookahead sit in shared memory at
// `entry`; no packet boundary lies within reach of a match starting in the group and no NUL precedes
// the group in its packet.  Start-anchored trie walk from every start position that fired.
__device__ __noinline__ void verify_simple(const union_params &p, const uint8_t *lut, const uint8_t *s_class,
                                           uint32_t *s_counts, const uint32_t *entry, uint32_t laneoff)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(entry);
    const uint32_t la = entry[4], g16 = entry[5];
    uint32_t S, cm = 0;
    S = LUT_AT(v.x, 0x5504) & 0x808080ffu;
    S = ((S << 8) | 0xffu) & LUT_AT(v.x, 0x5514);
    S = ((S << 8) | 0xffu) & LUT_AT(v.x, 0x5524);
    SV_STEP(v.x, 0x5534, 0);
    SV_STEP(v.y, 0x5504, 1); SV_STEP(v.y, 0x5514, 2); SV_STEP(v.y, 0x5524, 3); SV_STEP(v.y, 0x5534, 4);
    SV_STEP(v.z, 0x5504, 5); SV_STEP(v.z, 0x5514, 6); SV_STEP(v.z, 0x5524, 7); SV_STEP(v.z, 0x5534, 8);
    SV_STEP(v.w, 0x5504, 9); SV_STEP(v.w, 0x5514, 10); SV_STEP(v.w, 0x5524, 11); SV_STEP(v.w, 0x5534, 12);
    SV_STEP(la, 0x5504, 13); SV_STEP(la, 0x5514, 14); SV_STEP(la, 0x5524, 15);
    // starts at or after the group's first NUL are dead (serial.c:191)
    const uint32_t zm = zero_mask16(v);
    if (zm) cm &= (1u << (__ffs(zm) - 1)) - 1u;
    const uint8_t *eb = reinterpret_cast<const uint8_t *>(entry);
    const uint8_t *gb = p.bytes + 16ull * g16;
    while (cm) {
        const uint32_t i = __ffs(cm) - 1;
        cm &= cm - 1;
        uint32_t node = 0;
        for (uint32_t k = i;; k++) {
            const uint32_t c = k < 20 ? eb[k] : gb[k];
            const uint32_t e = __ldg(p.trie + node * p.n_class + s_class[c]);
            if (e == 0) break;
            node = e & 0x7fffffffu;
            if (e >> 31) count_hit(p, s_counts, __ldg(p.state_term + node));
        }
    }
}

// Slow path, complex entry: general walk.  g16 = group index (16-byte units from abs_base); the
// item holds packets [ks, ke); zones: 0 = before the item's first packet (dead), j = packet ks+j-1,
// > ke-ks = after the item's last packet (dead).  zone = the zone holding the group's first byte,
// dead = a NUL precedes it inside that packet.  Counts every pattern occurrence that starts in the
// group, lies inside one packet of the item and has no NUL before it in that packet.
__device__ __noinline__ void verify_complex(const union_params &p, const uint8_t *s_class, uint32_t *s_counts,
                                            uint32_t g16, uint32_t zone, bool dead, uint32_t ks, uint32_t ke)
{
    const uint64_t *off = p.offsets + ks;
    const uint32_t nbound = ke - ks;
    const uint64_t g = p.abs_base + 16ull * g16;
    uint64_t limit = g + 15 + p.max_len; // one past the last byte a match starting at g+15 can touch
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
        const uint32_t e = __ldg(p.next + state * p.n_class + s_class[c]);
        state = e & 0x7fffffffu;
        if (e >> 31) {
            const uint32_t o1 = __ldg(p.out_head + state + 1);
            for (uint32_t o = __ldg(p.out_head + state); o < o1; o++) {
                const uint32_t u = __ldg(p.out_id + o);
                if (pos + 1 - __ldg(p.uniq_len + u) < g + 16) // start >= g holds: the walk began at g in the root
                    count_hit(p, s_counts, u);
            }
        }
        if (pos >= g + 15 && state == 0) break; // no match in flight that started inside the group
    }
}

__global__ void __launch_bounds__(UN_THREADS, 1) kmpb_union_kernel(const __grid_constant__ union_params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *lut = smem;
    uint8_t *s_class = smem + UN_LUT_BYTES;
    uint32_t *qs_all = reinterpret_cast<uint32_t *>(smem + UN_LUT_BYTES + 256);
    uint32_t *qc_all = qs_all + UN_WARPS * UN_QCAP * UN_QS_WORDS;
    uint32_t *s_counts = p.counts_in_smem ? qc_all + UN_WARPS * UN_QCAP * UN_QC_WORDS : nullptr;

    for (uint32_t i = threadIdx.x; i < 256 * 32; i += UN_THREADS)
        reinterpret_cast<uint32_t *>(lut)[(i >> 5) * 64 + (i & 31)] = p.filter[i >> 5];
    for (uint32_t i = threadIdx.x; i < 256; i += UN_THREADS) s_class[i] = p.byte_class[i];
    if (s_counts)
        for (uint32_t i = threadIdx.x; i < p.n_uniq; i += UN_THREADS) s_counts[i] = 0;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t laneoff = lane << 2;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t *qs = qs_all + (threadIdx.x >> 5) * (UN_QCAP * UN_QS_WORDS);
    uint32_t *qc = qc_all + (threadIdx.x >> 5) * (UN_QCAP * UN_QC_WORDS);
    uint32_t qs_head = 0, qs_tail = 0, qc_head = 0, qc_tail = 0; // ring counters (warp-uniform)
    const uint32_t reach = 15u + p.max_len;

    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&p.work[0], 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= p.n_items) break;
        const uint32_t ks = p.items[item], ke = p.items[item + 1];
        if (ks >= ke) continue;
        const uint64_t b_abs = p.offsets[ks], e_abs = p.offsets[ke];
        if (b_abs == e_abs) continue;
        if (e_abs - b_abs >= (1ull << 31)) { // a packet over 2 GiB: outside the documented limits
            if (lane == 0) atomicOr(&p.work[1], 1u);
            continue;
        }
        const uint64_t row0 = b_abs & ~127ull;                 // absolute position of the item's first row
        const uint8_t *text = p.bytes + (row0 - p.abs_base);
        const uint64_t *off = p.offsets + ks;                  // item boundary j is off[j] - row0
        const uint32_t nbound = ke - ks;
        const uint32_t e_rel = (uint32_t)(e_abs - row0);
        const uint32_t load_end = (e_rel + 15u) & ~15u;
        const uint32_t g16_0 = (uint32_t)((row0 - p.abs_base) >> 4);

        // zone tracking (warp-uniform): zone 0 = before the first packet (dead), j = packet ks+j-1
        uint32_t zone = 0;
        uint32_t nb = (uint32_t)(b_abs - row0);        // boundary that ends the current zone
        uint32_t nb_next = (uint32_t)(off[1] - row0);  // the one after it (prefetched)
        bool dead = true;

        uint32_t g = lane * 16u;
        uint4 cur = make_uint4(0, 0, 0, 0), nx1 = make_uint4(0, 0, 0, 0);
        if (g < load_end) cur = __ldcs(reinterpret_cast<const uint4 *>(text + g));
        if (g + UN_ROW < load_end) nx1 = __ldcs(reinterpret_cast<const uint4 *>(text + g + UN_ROW));

        for (uint32_t row = 0; row < e_rel; row += UN_ROW, g += UN_ROW) {
            uint4 nx2 = make_uint4(0, 0, 0, 0);
            if (g + 2 * UN_ROW < load_end) nx2 = __ldcs(reinterpret_cast<const uint4 *>(text + g + 2 * UN_ROW));
            // 3 bytes of lookahead: first word of the next group (next lane, or lane 0 of the next row)
            const uint32_t la = __shfl_sync(FULL, lane == 0 ? nx1.x : cur.x, (lane + 1) & 31);

            // ---- shift-and filter over 19 bytes --------------------------------------------------
            uint32_t S, accA, accB, accC;
            S = LUT_AT(cur.x, 0x5504) & 0x808080ffu; // no history: only the NUL stage is pre-armed
            accA = S;
            SA_STEP(cur.x, 0x5514, accA);
            SA_STEP(cur.x, 0x5524, accA);
            accB = 0;
            SA_STEP(cur.x, 0x5534, accB);
            SA_WORD(cur.y, accB);
            SA_WORD(cur.z, accB);
            SA_WORD(cur.w, accB);
            accC = 0;
            SA_STEP(la, 0x5504, accC);
            SA_STEP(la, 0x5514, accC);
            SA_STEP(la, 0x5524, accC);
            const bool nul = ((accA | accB) >> 31) != 0;          // a NUL among my 16 bytes
            const bool cand = ((accB | accC) & 0x7f000000u) != 0; // a candidate start among my 16 positions
            const uint32_t nulm = __ballot_sync(FULL, nul);

            // ---- which packet am I in, and is it already dead? ----------------------------------
            const uint32_t row_end = row + UN_ROW;
            bool push_s = false, push_c = false, d0 = false;
            uint32_t kz = zone;
            if (nb >= row_end) { // no packet boundary inside this row
                d0 = dead || (nulm & lt) != 0;
                push_s = cand && !d0;
                if (push_s && nb < g + reach) { push_s = false; push_c = true; } // a boundary within reach
                dead = dead || nulm != 0;
            } else {
                const uint32_t zm = zero_mask16(cur);
                uint32_t lane_lo = 0;
                bool seg_dead = dead;
                for (;;) {
                    // the current zone covers lanes [lane_lo, lane_hi]: groups that START before nb
                    const int lane_hi = nb >= row_end ? 31 : ((int)(nb - row) - 1) >> 4;
                    if ((int)lane >= (int)lane_lo && (int)lane <= lane_hi) {
                        const uint32_t below = lane_lo >= 32 ? FULL : (1u << lane_lo) - 1u;
                        kz = zone;
                        d0 = seg_dead || (nulm & lt & ~below) != 0;
                        push_s = cand && !d0 && nb >= g + reach;
                        push_c = cand && !push_s && (!d0 || nb < g + 16u); // a boundary inside the group can revive it
                    }
                    if (nb >= row_end) {
                        const uint32_t below = lane_lo >= 32 ? FULL : (1u << lane_lo) - 1u;
                        dead = seg_dead || (nulm & ~below) != 0;
                        break;
                    }
                    const uint32_t lb = (nb - row) >> 4, ob = (nb - row) & 15u;
                    zone++;
                    if (zone > nbound) { // past the item's last packet
                        nb = UN_NOBOUND;
                        seg_dead = true;
                        lane_lo = lb + (ob ? 1u : 0u);
                        continue;
                    }
                    const uint32_t zlb = __shfl_sync(FULL, zm, lb);
                    if (ob) { lane_lo = lb + 1u; seg_dead = (zlb >> ob) != 0; }
                    else { lane_lo = lb; seg_dead = false; }
                    nb = nb_next;
                    nb_next = zone + 1 <= nbound ? (uint32_t)(off[zone + 1] - row0) : UN_NOBOUND;
                }
            }

            // ---- queue flagged groups; drain when a full warp's worth has gathered -------------
            const uint32_t ms = __ballot_sync(FULL, push_s);
            if (ms) {
                if (push_s) {
                    uint32_t *e = qs + ((qs_tail + __popc(ms & lt)) & (UN_QCAP - 1)) * UN_QS_WORDS;
                    *reinterpret_cast<uint4 *>(e) = cur;
                    *reinterpret_cast<uint2 *>(e + 4) = make_uint2(la, g16_0 + (g >> 4));
                }
                qs_tail += __popc(ms);
                __syncwarp();
                if (qs_tail - qs_head >= 32) {
                    verify_simple(p, lut, s_class, s_counts, qs + ((qs_head + lane) & (UN_QCAP - 1)) * UN_QS_WORDS, laneoff);
                    qs_head += 32;
                    __syncwarp();
                }
            }
            const uint32_t mc = __ballot_sync(FULL, push_c);
            if (mc) {
                if (push_c) {
                    uint32_t *e = qc + ((qc_tail + __popc(mc & lt)) & (UN_QCAP - 1)) * UN_QC_WORDS;
                    *reinterpret_cast<uint4 *>(e) = make_uint4(g16_0 + (g >> 4), kz | (d0 ? 0x80000000u : 0u), ks, ke);
                }
                qc_tail += __popc(mc);
                __syncwarp();
                if (qc_tail - qc_head >= 32) {
                    const uint4 e = *reinterpret_cast<const uint4 *>(qc + ((qc_head + lane) & (UN_QCAP - 1)) * UN_QC_WORDS);
                    qc_head += 32;
                    __syncwarp();
                    verify_complex(p, s_class, s_counts, e.x, e.y & 0x7fffffffu, (e.y >> 31) != 0, e.z, e.w);
                }
            }
            cur = nx1;
            nx1 = nx2;
        }
    }
    // leftovers
    __syncwarp();
    if (lane < qs_tail - qs_head)
        verify_simple(p, lut, s_class, s_counts, qs + ((qs_head + lane) & (UN_QCAP - 1)) * UN_QS_WORDS, laneoff);
    if (lane < qc_tail - qc_head) {
        const uint4 e = *reinterpret_cast<const uint4 *>(qc + ((qc_head + lane) & (UN_QCAP - 1)) * UN_QC_WORDS);
        verify_complex(p, s_class, s_counts, e.x, e.y & 0x7fffffffu, (e.y >> 31) != 0, e.z, e.w);
    }

    __syncthreads();
    if (s_counts)
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
    if ((b.abs_base & 511) || ((uintptr_t)b.d_bytes & 15))
        return kmpb_fail(KMPB_EINVAL, "payload buffer must be 16-byte aligned");
    if (b.end_byte - b.abs_base >= (1ull << 36))
        return kmpb_fail(KMPB_ELIMIT, "more than 64 GiB of payload in one batch");
    const uint64_t span = b.end_byte - b.first_byte;
    const uint32_t n_items = (uint32_t)((span + UN_ITEM_BYTES - 1) / UN_ITEM_BYTES);
    if ((size_t)n_items + 1 > ctx->items_cap) return kmpb_fail(KMPB_ESTATE, "union scratch too small");
    uint32_t *d_items = ctx->d_items + (size_t)slot * ctx->items_cap;
    uint32_t *d_work = ctx->d_work + slot * 4;

    const bool counts_in_smem = h.n_uniq <= KMPB_SMEM_COUNTS_MAX;
    const size_t smem = UN_SMEM_FIXED + (counts_in_smem ? (size_t)h.n_uniq * sizeof(uint32_t) : 0);
    if (!ctx->attr_union_set) {
        KMPB_CUDA(cudaFuncSetAttribute(kmpb_union_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(UN_SMEM_FIXED + KMPB_SMEM_COUNTS_MAX * 4)));
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
    p.max_len = h.max_len;
    p.counts_in_smem = counts_in_smem ? 1u : 0u;
    p.uniq_counts = (unsigned long long *)d_uniq_counts;
    const uint32_t warps_needed = n_items;
    int grid = (int)std::min<uint32_t>((uint32_t)ctx->sm_count, (warps_needed + UN_WARPS - 1) / UN_WARPS);
    if (grid < 1) grid = 1;
    if (ctx->profile) KMPB_CUDA(cudaEventRecord(ctx->ev_kernel[0], stream));
    kmpb_union_kernel<<<grid, UN_THREADS, smem, stream>>>(p);
    if (ctx->profile) KMPB_CUDA(cudaEventRecord(ctx->ev_kernel[1], stream));
    ctx->launches += 2;
    KMPB_CUDA(cudaGetLastError());
    return KMPB_OK;
}
