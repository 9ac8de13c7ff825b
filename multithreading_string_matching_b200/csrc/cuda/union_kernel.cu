// union_kernel.cu -- the union engine (KMPB_ENGINE_UNION): every payload byte is read from HBM once.
//
// Two levels.
//
//  FAST PATH (every byte).  A warp streams its work item -- a run of whole packets, ~128 KB of the
//  flat CSR byte buffer -- in rows of 512 contiguous bytes: one coalesced 16-byte load per lane, the
//  next row already in flight while the current one is scanned.  Each lane pushes its 16 bytes (+3
//  bytes of lookahead from its neighbour, by shuffle) through a 4-byte-deep shift-and filter over 8
//  buckets:  S = ((S << 8) | 0xff) & filter[byte].  filter[] lives in shared memory in a bank-private
//  layout (byte address = byte*256 + lane*4), so the one lookup per byte never bank-conflicts and
//  its address is a single PRMT.  Bits 24..30 of S say "the last 4 bytes are the first 4 bytes (or
//  all the bytes) of some pattern of bucket b"; bit 31 says "this byte is NUL".
//
//  SLOW PATH (rare).  Lanes whose 16 start positions raised a flag push (group, zone, dead?) into a
//  per-warp shared-memory queue.  When 32 entries have gathered, the warp drains them with every
//  lane busy: each lane walks the union automaton (the merged KMP DFAs, csrc/host/automaton.c) over
//  its group, honouring packet boundaries and the reference's "text ends at the first NUL" rule
//  (serial.c:191), and counts every pattern occurrence that STARTS inside the group.  Counts go to
//  shared-memory counters and leave the block as one atomic per distinct pattern.
//
//  Packet boundaries and NULs are tracked per warp while streaming: a work item starts and ends on
//  packet boundaries, so "was there a NUL earlier in this packet" is known from the ballots of the
//  rows already scanned.  No separators, no padding and no second pass over the payload.
#include <algorithm>

#include "kmpb_device.cuh"

constexpr int UN_THREADS = 1024; // one block per SM
constexpr int UN_WARPS = UN_THREADS / 32;
constexpr uint32_t UN_ROW = 512;              // bytes per warp row
constexpr uint32_t UN_ITEM_BYTES = 128 << 10; // target work-item size
constexpr uint32_t UN_QCAP = 64;              // queue entries per warp
constexpr uint32_t UN_LUT_BYTES = 256 * 256;  // 256-byte row per byte value; lanes use the first 128 B
constexpr uint32_t UN_NOBOUND = 0xffffffffu;
constexpr uint32_t FULL = 0xffffffffu;

struct union_params {
    const uint8_t *bytes; // device pointer to absolute byte abs_base (abs_base % 512 == 0)
    uint64_t abs_base;
    const uint64_t *offsets; // [n_packets+1], absolute
    uint32_t n_packets;
    const uint32_t *items; // [n_items+1] first packet of each work item
    uint32_t n_items;
    uint32_t *work; // [0] next item, [1] error flags
    const uint32_t *filter; // [256]
    const uint32_t *next;   // union DFA: next[state*n_class + class] = target | reports<<31
    const uint32_t *out_head;
    const uint32_t *out_id;
    const uint32_t *uniq_len;
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

struct item_ctx {
    const uint8_t *text;   // byte 0 of the item's first row
    const uint64_t *off;   // item boundary j is off[j] - row0
    uint64_t row0;         // absolute position of the item's first row
    uint32_t nbound;       // packets in the item; zones: 0 = before boundary 0 (dead), j = packet j-1,
                           // nbound+1 = after the last boundary (dead)
    uint32_t e_rel;        // end of the item's last packet, relative to row0
};

// Slow path for one flagged group.  g = group start relative to row0, zone = zone holding byte g,
// dead = a NUL precedes g inside that packet (always true in a dead zone).  Counts every pattern
// occurrence that starts in [g, g+16), lies inside one packet of this item and has no NUL before it
// in that packet.
__device__ __noinline__ void union_walk(const union_params &p, const item_ctx &it, const uint8_t *s_class,
                                        uint32_t *s_counts, uint32_t g, uint32_t zone, bool dead)
{
    uint32_t limit = g + 15 + p.max_len; // exclusive: one past the last byte a match starting at g+15 can touch
    if (limit > it.e_rel) limit = it.e_rel;
    uint32_t nb = zone <= it.nbound ? (uint32_t)(it.off[zone] - it.row0) : UN_NOBOUND;
    uint32_t state = 0;
    for (uint32_t pos = g; pos < limit; pos++) {
        while (pos == nb) { // crossing into the next packet (or out of the item)
            state = 0;
            dead = false;
            zone++;
            nb = zone <= it.nbound ? (uint32_t)(it.off[zone] - it.row0) : UN_NOBOUND;
        }
        if (zone > it.nbound) break;
        const uint32_t c = it.text[pos];
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
                const uint32_t start = pos + 1 - __ldg(p.uniq_len + u);
                if (start < g + 16) { // start >= g holds: the walk began at g in the root state
                    if (s_counts) atomicAdd(&s_counts[u], 1u);
                    else atomicAdd(&p.uniq_counts[u], 1ull);
                }
            }
        }
        if (pos >= g + 15 && state == 0) break; // no match in flight that started inside the group
    }
}

#define LUT_AT(word, sel) (*reinterpret_cast<const uint32_t *>(lut + __byte_perm((word), laneoff, (sel))))
#define SA_STEP(word, sel, acc)                  \
    do {                                         \
        const uint32_t m_ = LUT_AT(word, sel);   \
        S = ((S << 8) | 0xffu) & m_;             \
        acc |= S;                                \
    } while (0)
#define SA_WORD(word, acc)        \
    do {                          \
        SA_STEP(word, 0x5504, acc); \
        SA_STEP(word, 0x5514, acc); \
        SA_STEP(word, 0x5524, acc); \
        SA_STEP(word, 0x5534, acc); \
    } while (0)

__global__ void __launch_bounds__(UN_THREADS, 1) kmpb_union_kernel(const union_params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *lut = smem;
    uint8_t *s_class = smem + UN_LUT_BYTES;
    uint32_t *s_queue_all = reinterpret_cast<uint32_t *>(smem + UN_LUT_BYTES + 256);
    uint32_t *s_counts = p.counts_in_smem ? s_queue_all + UN_WARPS * UN_QCAP * 2 : nullptr;

    for (uint32_t i = threadIdx.x; i < 256 * 32; i += UN_THREADS)
        reinterpret_cast<uint32_t *>(lut)[(i >> 5) * 64 + (i & 31)] = p.filter[i >> 5];
    for (uint32_t i = threadIdx.x; i < 256; i += UN_THREADS) s_class[i] = p.byte_class[i];
    if (s_counts)
        for (uint32_t i = threadIdx.x; i < p.n_uniq; i += UN_THREADS) s_counts[i] = 0;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t laneoff = lane << 2;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t *queue = s_queue_all + (threadIdx.x >> 5) * (UN_QCAP * 2);

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
        item_ctx it;
        it.row0 = b_abs & ~127ull;
        it.text = p.bytes + (it.row0 - p.abs_base);
        it.off = p.offsets + ks;
        it.nbound = ke - ks;
        it.e_rel = (uint32_t)(e_abs - it.row0);
        const uint32_t load_end = (it.e_rel + 15u) & ~15u;

        // zone tracking (warp-uniform)
        uint32_t zone = 0;                                   // dead zone before the first packet
        uint32_t nb = (uint32_t)(b_abs - it.row0);           // boundary that ends the current zone
        uint32_t nb_next = (uint32_t)(it.off[1] - it.row0);  // the one after it (prefetched)
        bool dead = true;
        uint32_t qcount = 0;

        uint32_t g = lane * 16u;
        uint4 cur = make_uint4(0, 0, 0, 0);
        if (g < load_end) cur = __ldcs(reinterpret_cast<const uint4 *>(it.text + g));

        for (uint32_t row = 0; row < it.e_rel; row += UN_ROW, g += UN_ROW) {
            uint4 nxt = make_uint4(0, 0, 0, 0);
            if (g + UN_ROW < load_end) nxt = __ldcs(reinterpret_cast<const uint4 *>(it.text + g + UN_ROW));
            // 3 bytes of lookahead: first word of the next group (next lane, or lane 0 of the next row)
            const uint32_t la = __shfl_sync(FULL, lane == 0 ? nxt.x : cur.x, (lane + 1) & 31);

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
            const bool nul = ((accA | accB) >> 31) != 0;              // a NUL among my 16 bytes
            const bool cand = ((accB | accC) & 0x7f000000u) != 0;     // a candidate start among my 16 positions
            const uint32_t nulm = __ballot_sync(FULL, nul);

            // ---- which packet am I in, and is it already dead? ----------------------------------
            const uint32_t row_end = row + UN_ROW;
            bool push = false, d0 = false;
            uint32_t kz = zone;
            if (nb >= row_end) { // no packet boundary inside this row
                d0 = dead || (nulm & lt) != 0;
                push = cand && !d0;
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
                        push = cand && (!d0 || nb < g + 16u); // a boundary inside the group can revive it
                    }
                    if (nb >= row_end) {
                        const uint32_t below = lane_lo >= 32 ? FULL : (1u << lane_lo) - 1u;
                        dead = seg_dead || (nulm & ~below) != 0;
                        break;
                    }
                    const uint32_t lb = (nb - row) >> 4, ob = (nb - row) & 15u;
                    zone++;
                    if (zone > it.nbound) { // past the item's last packet
                        nb = UN_NOBOUND;
                        seg_dead = true;
                        lane_lo = lb + (ob ? 1u : 0u);
                        continue;
                    }
                    const uint32_t zlb = __shfl_sync(FULL, zm, lb);
                    if (ob) { lane_lo = lb + 1u; seg_dead = (zlb >> ob) != 0; }
                    else { lane_lo = lb; seg_dead = false; }
                    nb = nb_next;
                    nb_next = zone + 1 <= it.nbound ? (uint32_t)(it.off[zone + 1] - it.row0) : UN_NOBOUND;
                }
            }

            // ---- queue flagged groups; drain when a full warp's worth has gathered -------------
            const uint32_t pm = __ballot_sync(FULL, push);
            if (pm) {
                if (push) {
                    const uint32_t slot = qcount + __popc(pm & lt);
                    queue[2 * slot] = g;
                    queue[2 * slot + 1] = kz | (d0 ? 0x80000000u : 0u);
                }
                qcount += __popc(pm);
                __syncwarp();
                if (qcount >= 32) {
                    const uint32_t qg = queue[2 * lane], qk = queue[2 * lane + 1];
                    const uint32_t rest = qcount - 32;
                    uint32_t mg = 0, mk = 0;
                    if (lane < rest) { mg = queue[2 * (32 + lane)]; mk = queue[2 * (32 + lane) + 1]; }
                    __syncwarp();
                    if (lane < rest) { queue[2 * lane] = mg; queue[2 * lane + 1] = mk; }
                    qcount = rest;
                    __syncwarp();
                    union_walk(p, it, s_class, s_counts, qg, qk & 0x7fffffffu, (qk >> 31) != 0);
                }
            }
            cur = nxt;
        }
        // item ends: drain what is left (entries refer to this item's boundaries)
        if (qcount) {
            __syncwarp();
            if (lane < qcount) {
                const uint32_t qg = queue[2 * lane], qk = queue[2 * lane + 1];
                union_walk(p, it, s_class, s_counts, qg, qk & 0x7fffffffu, (qk >> 31) != 0);
            }
            __syncwarp();
        }
    }

    __syncthreads();
    if (s_counts)
        for (uint32_t u = threadIdx.x; u < p.n_uniq; u += UN_THREADS)
            if (s_counts[u]) atomicAdd(&p.uniq_counts[u], (unsigned long long)s_counts[u]);
}

// ---- host side ------------------------------------------------------------------------------------

static size_t union_smem_bytes(uint32_t n_uniq, bool *counts_in_smem)
{
    size_t base = UN_LUT_BYTES + 256 + (size_t)UN_WARPS * UN_QCAP * 2 * sizeof(uint32_t);
    *counts_in_smem = n_uniq <= KMPB_SMEM_COUNTS_MAX;
    return base + (*counts_in_smem ? (size_t)n_uniq * sizeof(uint32_t) : 0);
}

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
    const uint64_t span = b.end_byte - b.first_byte;
    const uint32_t n_items = (uint32_t)((span + UN_ITEM_BYTES - 1) / UN_ITEM_BYTES);
    if ((size_t)n_items + 1 > ctx->items_cap) return kmpb_fail(KMPB_ESTATE, "union scratch too small");
    uint32_t *d_items = ctx->d_items + (size_t)slot * ctx->items_cap;
    uint32_t *d_work = ctx->d_work + slot * 4;

    bool counts_in_smem;
    const size_t smem = union_smem_bytes(h.n_uniq, &counts_in_smem);
    if (!ctx->attr_union_set) {
        KMPB_CUDA(cudaFuncSetAttribute(kmpb_union_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(UN_LUT_BYTES + 256 + UN_WARPS * UN_QCAP * 8 + KMPB_SMEM_COUNTS_MAX * 4)));
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
