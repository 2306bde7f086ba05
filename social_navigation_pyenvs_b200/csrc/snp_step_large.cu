// snp_step_large.cu -- one very large crowd: shared-memory tiled all-pairs (N-body style) social force, then the same wall /
// desired / torque / Euler epilogue as the small-crowd kernel.  One sub-step on one stream:
//
//   k_tile_boxes    bounding box (+ max r+safety) of every 128-entity tile of the `others` view  [5][M] = x, y, vx, vy, r+s
//                   (first sub-step of a call only: afterwards the finish kernel writes the boxes of the tiles it produces)
//   pair phase      unit of work = (i-block of 128 or 256 agents, j-chunk of 128 entities): every agent accumulates the force of
//                   the chunk's entities in ascending j from a shared-memory tile and writes ONE partial sum per chunk.  A chunk
//                   is skipped when its box is farther from the i-block's box than the distance at which the pair law is
//                   identically zero in the arithmetic in use (exp underflow: r_i+s_i+r_j+s_j + 700 B in fp64, + 88 B in fp32
//                   with ex2.approx.ftz) -- an EXACT optimisation, the summed force is unchanged.
//                     all pairs:  k_large_pairs on the static (i-block, chunk) grid;
//                     culled:     k_large_cull lists the units in reach, persistent k_large_pairs_list CTAs work the list off.
//   k_large_finish  per agent: partial sums of the live chunks added in ascending chunk order (fixed, so the result does not depend
//                   on how the crowd is sharded over GPUs), goal switch, wall force, desired force, torque, Euler; writes the
//                   state in place and the agent's entry (and its tile's box) of the NEXT entity view -- on every rank.
//
// Small chunks matter once culling is on and the crowd is sharded: only a few percent of the units have anything to evaluate, and
// what one rank of an 8-way split keeps must still fill 148 SMs evenly (history: 4096-entity chunks 2.6 ms per sub-step at 65536
// humans and no gain from a second GPU; 512: 2.0 / 1.1 ms; 256 + exact culling on compact tiles + one agent per thread: 0.75 ms on
// one GPU, 0.158 on eight; 128 + the work list + grouped evaluation: 0.59 / 0.125).  The price is J = M/128 partial sums per LIVE chunk and agent.
// Reference: same as snp_step_small.cu (motion_model_manager.py:354-373,424-459; forces.py:63-151).
#include "snp_kernels.cuh"

namespace snp {
namespace {

constexpr int kTile = 128;           // entities per shared-memory tile == threads per block
// Register tiling of the pairs kernel: every staged entity is used for APT agents of the thread.  2 halves the shared-memory
// loads per pair (best when every ordered pair is evaluated: 7.8 vs 8.4 ms per sub-step at 65536 humans); 1 doubles the number of
// work units, which is what a culled fp64 step needs once the crowd is sharded.  The summation order per agent does not depend
// on it, so results stay bit-identical.
constexpr int kMaxAgentsPerThread = 2;
#ifndef SNP_LARGE_CHUNK
#define SNP_LARGE_CHUNK 128
#endif
// Culled steps (one agent per thread): entities evaluated together between two contact votes.  Measured in fp64 on one rank's slice
// of an 8-way split / on the whole 65536 crowd: 1 -> 0.1229 / 0.657 ms per sub-step, 2 -> 0.1167 / 0.621, 4 -> 0.1136 / 0.616,
// 8 -> 0.1106 / 0.604 (more registers, fewer resident CTAs per SM, but independent chains per warp when few warps are left on an SM).
#ifndef SNP_LARGE_GROUP
#define SNP_LARGE_GROUP 8
#endif
constexpr int kChunk = SNP_LARGE_CHUNK;  // entities per j-chunk (one partial sum each); fixed so results are sharding-independent

template <typename T> struct LargeArgs {
    KArgs<T> k;
    const T *others;
    long long M, self_offset;
    T *next_view;
    T *peers[8];  // next-view buffers of every rank (peer-mapped device pointers); the finish kernel stores into all of them
    int n_peers;
    T *partial;   // [J][2][N_local]
    T *boxes;     // [n_tiles][5] xmin, xmax, ymin, ymax, max(r+s)
    T *peer_boxes[8];  // snp_large_run_p2p: the NEXT view's tile boxes on every rank; the finish kernel writes its own tiles' boxes there
    unsigned char *live;  // [i-blocks][J]: 1 when the (i-block, chunk) CTA evaluated at least one tile and wrote its partial sums
    int J, n_tiles;
    int Jp;  // row stride of the `live` map: J rounded up to 16 (the finish kernel reads 16 flags per load)
    T cull_margin;  // distance beyond r+s sums at which the pair law is exactly zero; < 0 disables culling
    int boxes_ready;  // the tile boxes of `others` are already in `boxes` (written by the previous sub-step's producer)
    int apt;          // agents per thread of the pairs kernel (1 or 2): fixes the i-block size the `live` map is indexed by
    // culled steps: the (i-block, chunk) pairs that are near each other, listed by k_large_cull and handed out to persistent CTAs
    int *work;            // [i-blocks * J] entries i-block * J + chunk
    int *work_counters;   // [0] entries listed, [4 + q] entries handed out from queue q; zeroed by the finish kernel for the next sub-step
    int use_list;
    int *live_cnt, *live_list;  // per i-block: how many chunks are in reach, and which, ascending ([i-blocks], [i-blocks][Jp]) -- what the
                                // finish kernel walks instead of scanning the flags
};
constexpr int kMaxQueues = 256;                 // one queue of list entries per SM (entry k belongs to queue k mod n_queues)
constexpr int kWorkCounters = 4 + kMaxQueues;   // ints in front of the list

template <typename T> __device__ __forceinline__ T warp_min(T v) {
    for (int o = 16; o > 0; o >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    return v;
}
template <typename T> __device__ __forceinline__ T warp_max(T v) {
    for (int o = 16; o > 0; o >>= 1) { const T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}

// Block-wide (128 threads) box of per-thread (xmin, xmax, ymin, ymax, rsmax); result broadcast through `sbox`.
template <typename T> __device__ __forceinline__ void block_box(T xmin, T xmax, T ymin, T ymax, T rs, T *sbox /*[4][5]*/, T *out /*[5]*/) {
    xmin = warp_min(xmin); xmax = warp_max(xmax); ymin = warp_min(ymin); ymax = warp_max(ymax); rs = warp_max(rs);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sbox[w * 5 + 0] = xmin; sbox[w * 5 + 1] = xmax; sbox[w * 5 + 2] = ymin; sbox[w * 5 + 3] = ymax; sbox[w * 5 + 4] = rs; }
    __syncthreads();
    out[0] = min(min(sbox[0], sbox[5]), min(sbox[10], sbox[15]));
    out[1] = max(max(sbox[1], sbox[6]), max(sbox[11], sbox[16]));
    out[2] = min(min(sbox[2], sbox[7]), min(sbox[12], sbox[17]));
    out[3] = max(max(sbox[3], sbox[8]), max(sbox[13], sbox[18]));
    out[4] = max(max(sbox[4], sbox[9]), max(sbox[14], sbox[19]));
    __syncthreads();
}

template <typename T> __global__ void __launch_bounds__(kTile) k_tile_boxes(const T *others, long long M, T *boxes) {
    __shared__ T sbox[20];
    const long long j = (long long)blockIdx.x * kTile + threadIdx.x;
    const bool ok = j < M;
    const T big = Real<T>::inf();
    const T x = ok ? others[j] : T(0), y = ok ? others[M + j] : T(0), rs = ok ? others[4 * M + j] : T(0);
    T out[5];
    block_box<T>(ok ? x : big, ok ? x : -big, ok ? y : big, ok ? y : -big, rs, sbox, out);
    if (threadIdx.x < 5) boxes[(size_t)blockIdx.x * 5 + threadIdx.x] = out[threadIdx.x];
}

// The box of an i-block from the tile boxes alone: its agents are the entities of kAgentsPerThread consecutive tiles, whose union
// box contains them.
template <typename T> __device__ __forceinline__ void iblock_box(const LargeArgs<T> &la, int iblock, int apt, T *u /*[5]*/) {
    const long long t0 = (la.self_offset + (long long)iblock * apt * kTile) / kTile;
    const T inf = Real<T>::inf();
    u[0] = inf; u[1] = -inf; u[2] = inf; u[3] = -inf; u[4] = T(0);
    for (int q = 0; q < apt; ++q)
        if (t0 + q < la.n_tiles) {
            const T *b = la.boxes + (size_t)(t0 + q) * 5;
            u[0] = min(u[0], b[0]); u[1] = max(u[1], b[1]); u[2] = min(u[2], b[2]); u[3] = max(u[3], b[3]); u[4] = max(u[4], b[4]);
        }
}
// Is any tile of chunk `chunk` within reach of the box u?  (Beyond reach the pair law is identically zero.)
template <typename T> __device__ __forceinline__ bool chunk_near(const LargeArgs<T> &la, const T *u, int chunk) {
    const long long j_begin = (long long)chunk * kChunk, j_end = min(la.M, j_begin + kChunk);
    bool near = false;
    for (long long j0 = j_begin; j0 < j_end; j0 += kTile) {
        const T *b = la.boxes + (size_t)(j0 / kTile) * 5;
        const T gx = max(T(0), max(u[0] - b[1], b[0] - u[1]));
        const T gy = max(T(0), max(u[2] - b[3], b[2] - u[3]));
        const T reach = u[4] + b[4] + la.cull_margin;
        near |= !(fma_<T>(gx, gx, gy * gy) > reach * reach);
    }
    return near;
}

template <typename T> struct PairsSmem {
    alignas(16) unsigned char tile_raw[sizeof(Ent<T>) * kTile];
    T tile_rs[kTile];
    alignas(16) double exp_tbl_s[kExpN];
    T sbox[20];
    int item, steal;
};

// One (i-block, chunk) unit of the pair phase: the i-block's agents against the chunk's entities, one partial sum per agent.
template <typename T, int SOC, int kAgentsPerThread>
__device__ __forceinline__ void pairs_chunk(const LargeArgs<T> &la, PairsSmem<T> &sm, int iblock, int chunk, bool &have_tbl) {
    Ent<T> *tile = reinterpret_cast<Ent<T> *>(sm.tile_raw);
    T *tile_rs = sm.tile_rs;
    double *exp_tbl_s = sm.exp_tbl_s;
    T *sbox = sm.sbox;
    const KArgs<T> &a = la.k;
    const long long N = a.EN, M = la.M;
    const Params<T> &P = a.P;
    const long long j_begin = (long long)chunk * kChunk;
    const long long j_end = min(M, j_begin + kChunk);
    unsigned char *live_flag = la.live + (size_t)iblock * la.Jp + chunk;
    T mx[kAgentsPerThread], my[kAgentsPerThread], mvx[kAgentsPerThread], mvy[kAgentsPerThread], mrs[kAgentsPerThread];
    long long idx[kAgentsPerThread];
    T fsx[kAgentsPerThread], fsy[kAgentsPerThread];
    const T big = Real<T>::inf();
    T bx0 = big, bx1 = -big, by0 = big, by1 = -big, brs = T(0);
#pragma unroll
    for (int q = 0; q < kAgentsPerThread; ++q) {
        idx[q] = ((long long)iblock * kAgentsPerThread + q) * kTile + threadIdx.x;
        const bool live = idx[q] < N;
        const long long o = la.self_offset + (live ? idx[q] : 0);
        mx[q] = la.others[o]; my[q] = la.others[M + o]; mvx[q] = la.others[2 * M + o]; mvy[q] = la.others[3 * M + o]; mrs[q] = la.others[4 * M + o];
        fsx[q] = T(0); fsy[q] = T(0);
        if (live) { bx0 = min(bx0, mx[q]); bx1 = max(bx1, mx[q]); by0 = min(by0, my[q]); by1 = max(by1, my[q]); brs = max(brs, mrs[q]); }
    }
    T ibox[5];
    block_box<T>(bx0, bx1, by0, by1, brs, sbox, ibox);

    const bool sym = a.symmetric != 0;
    bool any_tile = false;
    for (long long j0 = j_begin; j0 < j_end; j0 += kTile) {
        if (la.cull_margin >= T(0)) {  // uniform across the CTA
            const T *b = la.boxes + (size_t)(j0 / kTile) * 5;
            const T gx = max(T(0), max(ibox[0] - b[1], b[0] - ibox[1]));
            const T gy = max(T(0), max(ibox[2] - b[3], b[2] - ibox[3]));
            const T reach = ibox[4] + b[4] + la.cull_margin;
            if (fma_<T>(gx, gx, gy * gy) > reach * reach) continue;
        }
        if (sizeof(T) == 8 && !have_tbl) { exp_table_init(exp_tbl_s); have_tbl = true; }
        any_tile = true;
        __syncthreads();
        const long long j = j0 + threadIdx.x;
        if (j < j_end) {
            tile[threadIdx.x] = Ent<T>{la.others[j], la.others[M + j], la.others[2 * M + j], la.others[3 * M + j]};
            tile_rs[threadIdx.x] = la.others[4 * M + j];
        }
        __syncthreads();
        const int cnt = (int)min((long long)kTile, j_end - j0);
        int t_first = 0;
#if SNP_LARGE_GROUP > 1
        if constexpr (kAgentsPerThread == 1 && SOC != 2) {
            // one agent per thread (culled steps): SNP_LARGE_GROUP consecutive entities are evaluated branch-free and share ONE contact vote
            // (independent chains in flight when few warps are left on the SM); added in entity order, so the sum is unchanged
            for (; t_first + SNP_LARGE_GROUP <= cnt; t_first += SNP_LARGE_GROUP) {
                T fx[SNP_LARGE_GROUP], fy[SNP_LARGE_GROUP];
                bool contact = false;
#pragma unroll
                for (int u = 0; u < SNP_LARGE_GROUP; ++u) {
                    const Ent<T> o = tile[t_first + u];
                    contact |= Real<T>::positive_(pair_eval<T, SOC, false>(P, exp_tbl_s, mx[0], my[0], mvx[0], mvy[0], mrs[0], o.x, o.y, o.vx, o.vy, tile_rs[t_first + u], fx[u], fy[u]));
                }
                if (__any_sync(0xffffffffu, contact)) {
#pragma unroll
                    for (int u = 0; u < SNP_LARGE_GROUP; ++u) {
                        const Ent<T> o = tile[t_first + u];
                        pair_eval<T, SOC, true>(P, exp_tbl_s, mx[0], my[0], mvx[0], mvy[0], mrs[0], o.x, o.y, o.vx, o.vy, tile_rs[t_first + u], fx[u], fy[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < SNP_LARGE_GROUP; ++u) { fsx[0] += fx[u]; fsy[0] += fy[u]; }
            }
        }
#endif
#pragma unroll 2
        for (int t = t_first; t < cnt; ++t) {
            const Ent<T> o = tile[t];
            const T rsj = tile_rs[t];
            const long long jj = j0 + t - la.self_offset;  // index of the entity in this crowd's numbering
            // the self pair (jj == idx[q]) contributes exactly zero by construction (tiny_ in pair_eval)
            if constexpr (sizeof(T) == 4) {
                // the evaluations of the thread's agents are independent and branch-free; ONE vote covers the rare contact
                // re-evaluation of all of them
                T fx[kAgentsPerThread], fy[kAgentsPerThread];
                bool contact = false;
#pragma unroll
                for (int q = 0; q < kAgentsPerThread; ++q) {
                    const bool sw = SOC == 2 && sym && jj < idx[q];
                    if (SOC == 2) {
                        contact |= pair_eval<T, SOC, false>(P, exp_tbl_s, sw ? o.x : mx[q], sw ? o.y : my[q], sw ? o.vx : mvx[q], sw ? o.vy : mvy[q], sw ? rsj : mrs[q],
                                                            sw ? mx[q] : o.x, sw ? my[q] : o.y, sw ? mvx[q] : o.vx, sw ? mvy[q] : o.vy, sw ? mrs[q] : rsj, fx[q], fy[q]) > T(0);
                        fx[q] = sw ? -fx[q] : fx[q]; fy[q] = sw ? -fy[q] : fy[q];
                    } else {
                        contact |= pair_eval<T, SOC, false>(P, exp_tbl_s, mx[q], my[q], mvx[q], mvy[q], mrs[q], o.x, o.y, o.vx, o.vy, rsj, fx[q], fy[q]) > T(0);
                    }
                }
                if (__any_sync(0xffffffffu, contact)) {
#pragma unroll
                    for (int q = 0; q < kAgentsPerThread; ++q) {
                        const bool sw = SOC == 2 && sym && jj < idx[q];
                        if (SOC == 2) {
                            pair_eval<T, SOC, true>(P, exp_tbl_s, sw ? o.x : mx[q], sw ? o.y : my[q], sw ? o.vx : mvx[q], sw ? o.vy : mvy[q], sw ? rsj : mrs[q],
                                                    sw ? mx[q] : o.x, sw ? my[q] : o.y, sw ? mvx[q] : o.vx, sw ? mvy[q] : o.vy, sw ? mrs[q] : rsj, fx[q], fy[q]);
                            fx[q] = sw ? -fx[q] : fx[q]; fy[q] = sw ? -fy[q] : fy[q];
                        } else {
                            pair_eval<T, SOC, true>(P, exp_tbl_s, mx[q], my[q], mvx[q], mvy[q], mrs[q], o.x, o.y, o.vx, o.vy, rsj, fx[q], fy[q]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < kAgentsPerThread; ++q) { fsx[q] += fx[q]; fsy[q] += fy[q]; }
            } else {  // fp64, two agents per thread (all pairs) or the tail of a tile: one evaluation at a time
#pragma unroll
                for (int q = 0; q < kAgentsPerThread; ++q) {
                    T fx, fy;
                    if (SOC == 2) {
                        const bool sw = sym && jj < idx[q];
                        pair_force<T, SOC>(P, exp_tbl_s, 0xffffffffu, sw ? o.x : mx[q], sw ? o.y : my[q], sw ? o.vx : mvx[q], sw ? o.vy : mvy[q], sw ? rsj : mrs[q],
                                           sw ? mx[q] : o.x, sw ? my[q] : o.y, sw ? mvx[q] : o.vx, sw ? mvy[q] : o.vy, sw ? mrs[q] : rsj, fx, fy);
                        fx = sw ? -fx : fx; fy = sw ? -fy : fy;
                    } else {
                        pair_force<T, SOC>(P, exp_tbl_s, 0xffffffffu, mx[q], my[q], mvx[q], mvy[q], mrs[q], o.x, o.y, o.vx, o.vy, rsj, fx, fy);
                    }
                    fsx[q] += fx; fsy[q] += fy;
                }
            }
        }
    }
    if (threadIdx.x == 0) *live_flag = any_tile ? 1 : 0;
    if (!any_tile) return;  // every tile was culled: the chunk contributes exactly zero and the finish kernel skips it
#pragma unroll
    for (int q = 0; q < kAgentsPerThread; ++q)
        if (idx[q] < N) {
            la.partial[((size_t)chunk * 2 + 0) * N + idx[q]] = fsx[q];
            la.partial[((size_t)chunk * 2 + 1) * N + idx[q]] = fsy[q];
        }
}

// Regular grid (i-block, chunk): every ordered pair, or -- with culling -- CTAs that first decide from the tile boxes whether their
// chunk is in reach at all, before any state is loaded or any barrier is reached.  A culled CTA leaves its partial sums unwritten
// and says so in `live`.
template <typename T, int SOC, int kAgentsPerThread>
__global__ void __launch_bounds__(kTile) k_large_pairs(const LargeArgs<T> la) {
    __shared__ PairsSmem<T> sm;
    if (la.cull_margin >= T(0) && la.self_offset % kTile == 0) {
        T u[5];
        iblock_box<T>(la, blockIdx.x, kAgentsPerThread, u);
        if (!chunk_near<T>(la, u, blockIdx.y)) {  // uniform across the CTA
            if (threadIdx.x == 0) la.live[(size_t)blockIdx.x * la.Jp + blockIdx.y] = 0;
            return;
        }
    }
    bool have_tbl = false;  // the 16 kB exp table is staged only by CTAs that evaluate at least one tile (most are culled)
    pairs_chunk<T, SOC, kAgentsPerThread>(la, sm, blockIdx.x, blockIdx.y, have_tbl);
}

// Culled steps of a wide crowd: of the i-blocks x J grid only a few percent of the CTAs have anything to evaluate (65536 humans on
// 512 m: 12 800 of 131 072), and what is left is too uneven for a static grid -- a rank of an 8-way split ends with 1 600 CTAs
// for 888 resident slots.  So the reach test runs first, one thread per (i-block, chunk), and lists the pairs that are near;
// persistent CTAs (one per resident slot, exp table staged once) then work the list off.  WHICH CTA takes which entry matters: with
// one shared counter the ~400 entries left when every CTA has had its first land on the SMs at random -- some get 7, some none --
// and the slowest SM sets the time (measured: 45 % over the ideal).  So entry k belongs to queue k mod n_SMs, a CTA serves the
// queue of the SM it runs on (%smid) and only when that is empty looks for another queue with entries left: every SM gets the
// same number of entries whatever the placement of the CTAs, and every entry is taken whichever CTAs exist.  Partial sums land
// where the grid version puts them: results are bit-identical.
template <typename T> __global__ void __launch_bounds__(256) k_large_cull(const LargeArgs<T> la) {
    __shared__ int warp_cnt[8];
    T u[5];
    iblock_box<T>(la, blockIdx.x, la.apt, u);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *mine = la.live_list + (size_t)blockIdx.x * la.Jp;
    int listed = 0;  // chunks of this i-block listed so far (uniform)
    for (int c0 = 0; c0 < la.J; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        const bool near = c < la.J && chunk_near<T>(la, u, c);
        if (c < la.J) la.live[(size_t)blockIdx.x * la.Jp + c] = near ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, near);
        const int below = __popc(m & ((1u << lane) - 1u));
        if (m) {  // the global list of work units (any order)
            int base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(la.work_counters, __popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (near) la.work[base + below] = blockIdx.x * la.J + c;
        }
        // this i-block's own list, ascending: the order the finish kernel adds the partial sums in
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int v = warp_cnt[w]; before += w < warp ? v : 0; total += v; }
        if (near) mine[listed + before + below] = c;
        listed += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) la.live_cnt[blockIdx.x] = listed;
}

template <typename T, int SOC, int kAgentsPerThread>
__global__ void __launch_bounds__(kTile) k_large_pairs_list(const LargeArgs<T> la, const int n_queues) {
    __shared__ PairsSmem<T> sm;
    bool have_tbl = false;
    const int n_items = la.work_counters[0];
    int *cursor = la.work_counters + 4;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    int q = (int)(smid % (unsigned)n_queues);
    for (;;) {
        __syncthreads();  // the previous unit is done with the tile, the boxes and sm.item
        if (threadIdx.x == 0) {
            const long long k = (long long)q + (long long)atomicAdd(cursor + q, 1) * n_queues;
            sm.item = k < n_items ? (int)k : -1;
            sm.steal = 0x7fffffff;
        }
        __syncthreads();
        const int item = sm.item;
        if (item >= 0) {
            const int w = la.work[item];
            pairs_chunk<T, SOC, kAgentsPerThread>(la, sm, w / la.J, w % la.J, have_tbl);
            continue;
        }
        // this queue is empty (and stays so: the counters only grow): move to the nearest queue that still has entries, if any
        for (int t = threadIdx.x; t < n_queues; t += kTile) {
            const int c = *reinterpret_cast<volatile int *>(cursor + t);
            if ((long long)t + (long long)c * n_queues < n_items) atomicMin(&sm.steal, (t - q + n_queues) % n_queues);
        }
        __syncthreads();
        if (sm.steal == 0x7fffffff) return;
        q = (q + sm.steal) % n_queues;
    }
}

template <typename T, int OBS, int HEADED>
__global__ void __launch_bounds__(kTile) k_large_finish(const LargeArgs<T> la) {
    using R = Real<T>;
    const KArgs<T> &a = la.k;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nseg = a.W * a.S;
    double *exp_tbl_s = reinterpret_cast<double *>(smem_raw);
    constexpr size_t kTbl = sizeof(double) * kExpN;
    Seg<T> *segs = reinterpret_cast<Seg<T> *>(smem_raw + kTbl);
    int *seg_cnt = reinterpret_cast<int *>(smem_raw + kTbl + ((sizeof(Seg<T>) * (size_t)nseg + 15) & ~size_t(15)));
    if (sizeof(T) == 8 && nseg > 0) exp_table_init(exp_tbl_s);  // only the wall force evaluates exp here
    for (int k = threadIdx.x; k < nseg; k += blockDim.x) {
        const T *w = a.walls + (size_t)k * 4;
        segs[k] = make_seg<T>(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < a.W; k += blockDim.x) {
        int c = 0;
        while (c < a.S && segs[(size_t)k * a.S + c].ax == segs[(size_t)k * a.S + c].ax) ++c;
        seg_cnt[k] = c;
    }
    __syncthreads();
    const long long N = a.EN, M = la.M;
    const long long i = (long long)blockIdx.x * kTile + threadIdx.x;
    if (la.use_list && blockIdx.x == 0)  // the pair phase is over: empty list and queues for the next sub-step
        for (int k = threadIdx.x; k < kWorkCounters; k += kTile) la.work_counters[k] = 0;
    const bool fold_boxes = la.peer_boxes[0] != nullptr;  // uniform; only offered when N is a whole number of tiles, so that every
    if (i >= N) return;                                    // thread of every block reaches the box reduction at the end
    const unsigned vote_mask = __activemask();  // the lanes that own an agent (the tail warp is partial)
    const Params<T> &P = a.P;
    Agent<T> m;
    m.px = a.dyn[SNP_DYN_PX * N + i]; m.py = a.dyn[SNP_DYN_PY * N + i];
    m.vx = a.dyn[SNP_DYN_VX * N + i]; m.vy = a.dyn[SNP_DYN_VY * N + i];
    m.dfx = a.dyn[SNP_DYN_DFX * N + i]; m.dfy = a.dyn[SNP_DYN_DFY * N + i];
    if (HEADED) {
        m.th = a.dyn[SNP_DYN_TH * N + i]; m.bvx = a.dyn[SNP_DYN_BVX * N + i]; m.bvy = a.dyn[SNP_DYN_BVY * N + i];
        m.om = a.dyn[SNP_DYN_OM * N + i];
        R::sincos_(m.th, &m.sn, &m.cs);
        m.vx = np_mv(m.cs, -m.sn, m.bvx, m.bvy);  // mmm:448; the same value k_large_publish / the previous step put in the view
        m.vy = np_mv(m.sn, m.cs, m.bvx, m.bvy);
    } else { m.th = m.bvx = m.bvy = m.om = T(0); m.cs = T(1); m.sn = T(0); }
    m.r = a.stat[SNP_STAT_R * N + i]; m.m = a.stat[SNP_STAT_M * N + i]; m.vd = a.stat[SNP_STAT_VD * N + i];
    m.rs = m.r + a.stat[SNP_STAT_SAFETY * N + i];
    agent_static<T>(P, m);
    int gidx = a.goal_idx[i];
    const int gcnt = a.goal_cnt[i];
    m.gx = a.goals[((size_t)gidx * 2 + 0) * N + i]; m.gy = a.goals[((size_t)gidx * 2 + 1) * N + i];
    T fsx = T(0), fsy = T(0);
    // Sum of the live chunks' partial sums in ascending chunk order (the order is what makes the result independent of the
    // sharding).  The flags are the same for the whole CTA and live chunks come in runs; they are read 64 at a time, and for every
    // 16 flags with a live one the loads of all live partial sums are issued together before the (ordered) additions -- a
    // flag-by-flag loop serialised one L2 round trip per live chunk.
    const uint4 *lv = reinterpret_cast<const uint4 *>(la.live + (size_t)(i / (kTile * la.apt)) * la.Jp);
    const int n_words = la.use_list ? 0 : la.Jp >> 4;
    if (la.use_list) {
        // the cull kernel left the i-block's live chunks as an ascending list: sixteen indices per step, the next sixteen already on
        // their way while the partial sums of the current ones are loaded
        const int4 *ll = reinterpret_cast<const int4 *>(la.live_list + (size_t)(i / (kTile * la.apt)) * la.Jp);
        const int cnt = la.live_cnt[i / (kTile * la.apt)];
        int4 nxt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) nxt[u] = ll[u];  // (a row is Jp >= 16 entries long; entries beyond cnt are ignored)
        for (int k = 0; k < cnt; k += 16) {
            int4 cur[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) cur[u] = nxt[u];
            if (k + 16 < cnt) {
#pragma unroll
                for (int u = 0; u < 4; ++u) nxt[u] = ll[(k >> 2) + 4 + u];
            }
            const int idx[16] = {cur[0].x, cur[0].y, cur[0].z, cur[0].w, cur[1].x, cur[1].y, cur[1].z, cur[1].w,
                                 cur[2].x, cur[2].y, cur[2].z, cur[2].w, cur[3].x, cur[3].y, cur[3].z, cur[3].w};
            T vx[16], vy[16];
#pragma unroll
            for (int b = 0; b < 16; ++b)
                if (k + b < cnt) { vx[b] = la.partial[((size_t)idx[b] * 2 + 0) * N + i]; vy[b] = la.partial[((size_t)idx[b] * 2 + 1) * N + i]; }
#pragma unroll
            for (int b = 0; b < 16; ++b)
                if (k + b < cnt) { fsx += vx[b]; fsy += vy[b]; }
        }
    }
    for (int w0 = 0; w0 < n_words; w0 += 4) {
        uint4 f[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) f[u] = (w0 + u < n_words) ? lv[w0 + u] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (!(f[u].x | f[u].y | f[u].z | f[u].w)) continue;
            const unsigned word[4] = {f[u].x, f[u].y, f[u].z, f[u].w};  // sixteen flags, one per byte
            const int p0 = (w0 + u) << 4;
            T vx[16], vy[16];
            bool on[16];
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                on[b] = ((word[b >> 2] >> (8 * (b & 3))) & 0xffu) != 0 && p0 + b < la.J;  // (bytes beyond J are padding nobody wrote)
                if (on[b]) { vx[b] = la.partial[((size_t)(p0 + b) * 2 + 0) * N + i]; vy[b] = la.partial[((size_t)(p0 + b) * 2 + 1) * N + i]; }
            }
#pragma unroll
            for (int b = 0; b < 16; ++b)
                if (on[b]) { fsx += vx[b]; fsy += vy[b]; }
        }
    }

    const T dg = np_norm(m.gx - m.px, m.gy - m.py);
    if (a.numba ? (dg <= m.r) : (dg < m.r)) {
        gidx = (gidx + 1 >= gcnt) ? 0 : gidx + 1;
        m.gx = a.goals[((size_t)gidx * 2 + 0) * N + i]; m.gy = a.goals[((size_t)gidx * 2 + 1) * N + i];
    }
    T fox = T(0), foy = T(0);
    if (a.W > 0) obstacle_force<T, OBS>(P, exp_tbl_s, vote_mask, segs, seg_cnt, a.W, a.S, a.numba != 0, m.px, m.py, m.vx, m.vy, m.rs, fox, foy);
    desired_force<T>(P, m, a.numba != 0);
    integrate<T, HEADED>(P, m, fox, foy, fsx, fsy, a.dt);
    a.dyn[SNP_DYN_PX * N + i] = m.px; a.dyn[SNP_DYN_PY * N + i] = m.py;
    a.dyn[SNP_DYN_VX * N + i] = m.vx; a.dyn[SNP_DYN_VY * N + i] = m.vy;
    a.dyn[SNP_DYN_DFX * N + i] = m.dfx; a.dyn[SNP_DYN_DFY * N + i] = m.dfy;
    if (HEADED) {
        a.dyn[SNP_DYN_TH * N + i] = m.th; a.dyn[SNP_DYN_BVX * N + i] = m.bvx; a.dyn[SNP_DYN_BVY * N + i] = m.bvy;
        a.dyn[SNP_DYN_OM * N + i] = m.om;
    }
    a.goal_idx[i] = gidx;
    if (la.n_peers > 0) {
        // the all-gather of the entity view fused into the producer: every rank's copy of the NEXT view receives this agent's
        // entry through peer (NVLink) stores, so no separate collective runs between sub-steps -- only a barrier
        const long long o = la.self_offset + i;
        for (int p = 0; p < la.n_peers; ++p) {
            T *v = la.peers[p];
            v[o] = m.px; v[M + o] = m.py; v[2 * M + o] = m.vx; v[3 * M + o] = m.vy; v[4 * M + o] = m.rs;
        }
    } else if (la.next_view) {
        const long long o = la.self_offset + i;
        la.next_view[o] = m.px; la.next_view[M + o] = m.py; la.next_view[2 * M + o] = m.vx; la.next_view[3 * M + o] = m.vy;
        la.next_view[4 * M + o] = m.rs;
    }
    if (fold_boxes) {
        // k_tile_boxes folded into the producer: this block's agents ARE one tile of the next view (self_offset is a multiple of
        // the tile size), so its bounding box goes to every rank's next-box table with the entries themselves
        __shared__ T sbox2[20];
        T out[5];
        block_box<T>(m.px, m.px, m.py, m.py, m.rs, sbox2, out);
        if (threadIdx.x < 5) {
            const long long tile = la.self_offset / kTile + blockIdx.x;
            for (int p = 0; p < la.n_peers; ++p) la.peer_boxes[p][tile * 5 + threadIdx.x] = out[threadIdx.x];
        }
    }
}

// Cross-rank barrier of our own between sub-steps (one tiny kernel instead of a host-driven collective): slot [r] of every rank's
// flag array receives rank r's epoch through a peer store; a rank passes once all its slots have reached the epoch.  Epochs only
// grow, so nothing is ever reset and a fast rank that is already signalling the next barrier cannot release a slow one early.
// The kernel boundary orders the finish kernel's peer stores before the signal (fence.sc.sys + release store); the acquire loads
// order the next sub-step's reads after it.  A rank that never arrives trips the time-out instead of hanging the GPU.
struct BarrierArgs {
    unsigned long long *flags[8];  // every rank's flag array [world] (peer-mapped), index = rank
    int world, rank;
    unsigned long long epoch;
    int *error;
};

__global__ void k_rank_barrier(const BarrierArgs b) {
    const int t = threadIdx.x;
    if (t < b.world) {
        __threadfence_system();
        unsigned long long *dst = b.flags[t] + b.rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(b.epoch) : "memory");
        const unsigned long long *src = b.flags[b.rank] + t;
        const long long t0 = clock64();
        unsigned long long seen = 0;
        while (true) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
            if (seen >= b.epoch) break;
            if (clock64() - t0 > 40000000000LL) { *b.error = 1; break; }  // ~20 s: a peer died; report instead of spinning forever
        }
    }
    __syncthreads();
    __threadfence_system();
}

// x, y, vx, vy, r+safety of every agent into a [5][stride] view (v = R(yaw) bv for headed models, mmm:448).
template <typename T> __global__ void k_large_publish(const T *dyn, const T *stat, long long N, int headed, T *view, long long stride, long long offset) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    T vx = dyn[SNP_DYN_VX * N + i], vy = dyn[SNP_DYN_VY * N + i];
    if (headed) {
        T s, c;
        Real<T>::sincos_(dyn[SNP_DYN_TH * N + i], &s, &c);
        const T bx = dyn[SNP_DYN_BVX * N + i], by = dyn[SNP_DYN_BVY * N + i];
        vx = np_mv(c, -s, bx, by); vy = np_mv(s, c, bx, by);
    }
    const long long o = offset + i;
    view[o] = dyn[SNP_DYN_PX * N + i]; view[stride + o] = dyn[SNP_DYN_PY * N + i];
    view[2 * stride + o] = vx; view[3 * stride + o] = vy;
    view[4 * stride + o] = stat[SNP_STAT_R * N + i] + stat[SNP_STAT_SAFETY * N + i];
}

template <typename T, int SOC, int OBS, int HEADED> int launch_large(const LargeArgs<T> &la, cudaStream_t st) {
    const long long N = la.k.EN;
    const int nseg = la.k.W * la.k.S;
    if (sizeof(T) == 8) SNP_CUDA_OK(ensure_exp_table());
    if (!la.boxes_ready) k_tile_boxes<T><<<(unsigned)la.n_tiles, kTile, 0, st>>>(la.others, la.M, la.boxes);
    const long long per_block = (long long)kTile * la.apt;
    dim3 grid((unsigned)((N + per_block - 1) / per_block), (unsigned)la.J);
    if (la.use_list) {
        if (!la.boxes_ready) SNP_CUDA_OK(cudaMemsetAsync(la.work_counters, 0, kWorkCounters * sizeof(int), st));  // first sub-step of a call; later ones: the finish kernel
        k_large_cull<T><<<grid.x, 256, 0, st>>>(la);
        static int slots1 = 0, slots2 = 0;  // resident CTAs per SM of the two instantiations
        const int n_queues = device_sm_count() < kMaxQueues ? device_sm_count() : kMaxQueues;
        if (la.apt == 1) {
            if (!slots1) SNP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&slots1, k_large_pairs_list<T, SOC, 1>, kTile, 0));
            k_large_pairs_list<T, SOC, 1><<<(unsigned)(device_sm_count() * (slots1 > 0 ? slots1 : 1)), kTile, 0, st>>>(la, n_queues);
        } else {
            if (!slots2) SNP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&slots2, k_large_pairs_list<T, SOC, 2>, kTile, 0));
            k_large_pairs_list<T, SOC, 2><<<(unsigned)(device_sm_count() * (slots2 > 0 ? slots2 : 1)), kTile, 0, st>>>(la, n_queues);
        }
        count_launch();
    } else if (la.apt == 1) k_large_pairs<T, SOC, 1><<<grid, kTile, 0, st>>>(la);
    else k_large_pairs<T, SOC, 2><<<grid, kTile, 0, st>>>(la);
    const size_t smem = sizeof(double) * kExpN + ((sizeof(Seg<T>) * (size_t)nseg + 15) & ~size_t(15)) + sizeof(int) * (la.k.W + 1) + 16;
    auto fin = k_large_finish<T, OBS, HEADED>;
    if (smem > 48 * 1024) SNP_CUDA_OK(cudaFuncSetAttribute(fin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fin<<<(unsigned)((N + kTile - 1) / kTile), kTile, smem, st>>>(la);
    count_launch(la.boxes_ready ? 2 : 3);
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

inline long long large_J(long long M) { return (M + kChunk - 1) / kChunk; }
inline long long large_tiles(long long M) { return (M + kTile - 1) / kTile; }
inline long long large_iblocks(long long N) { return (N + kTile - 1) / kTile; }  // sized for one agent per thread (the finer split)

template <typename T> int run_large(const snp_crowd *c, const snp_step_opts *o, const void *others, long long M, long long self_offset,
                                    void *next_view, const void *const *peer_views, int n_peers, void *scratch, long long scratch_bytes,
                                    cudaStream_t st, const void *cur_boxes = nullptr, bool boxes_ready = false, const void *const *peer_next_boxes = nullptr) {
    LargeArgs<T> la;
    la.n_peers = n_peers;
    for (int p = 0; p < 8; ++p) la.peers[p] = (p < n_peers) ? (T *)peer_views[p] : nullptr;
    for (int p = 0; p < 8; ++p) la.peer_boxes[p] = (peer_next_boxes && p < n_peers) ? (T *)peer_next_boxes[p] : nullptr;
    la.boxes_ready = boxes_ready ? 1 : 0;
    KArgs<T> &a = la.k;
    a.E = 1; a.N = 0; a.G = c->G; a.EN = (long long)c->E * c->N;
    a.dyn = (T *)c->dyn; a.stat = (const T *)c->stat; a.goals = (const T *)c->goals; a.goal_idx = c->goal_idx; a.goal_cnt = c->goal_cnt;
    a.agent_params = nullptr; a.P = make_params<T>(c->params); a.robot = nullptr;
    a.walls = (const T *)c->walls; a.W = c->W; a.S = c->W > 0 ? c->S : 0; a.walls_per_env = 0;
    a.consider_robot = 0; a.symmetric = o->symmetric; a.numba = o->numba_compat; a.n_substeps = 1; a.robot_mode = 0;
    a.dt = (T)o->dt; a.dt_d = o->dt; a.action = nullptr; a.pre_checks = a.post_checks = a.track_touch = 0;
    a.time_now = nullptr; a.flags = nullptr; a.checks = nullptr; a.epw = 1; a.gpb = 1; a.mapping = 0; a.full_pair_loop = 1; a.respawn = 0; a.robot_type = 0; a.RP = a.P; a.robot_every = 0; a.robot_phase = 0; a.robot_dt = T(0);
    la.others = (const T *)others; la.M = M; la.self_offset = self_offset; la.next_view = (T *)next_view;
    la.J = (int)large_J(M); la.n_tiles = (int)large_tiles(M);
    la.Jp = (la.J + 15) & ~15;
    const long long live_bytes = large_iblocks(a.EN) * la.Jp, list_entries = large_iblocks(a.EN) * la.J;
    const long long need = (long long)sizeof(T) * ((long long)la.J * 2 * a.EN + (long long)la.n_tiles * 5) + live_bytes + 64 +
                           4 * (kWorkCounters + list_entries + ((large_iblocks(a.EN) + 3) & ~3LL) + large_iblocks(a.EN) * la.Jp);
    if (!scratch || scratch_bytes < need) { set_error("snp_large_step: scratch of %lld bytes needed, %lld given", need, scratch_bytes); return SNP_ERR_INVALID; }
    la.partial = (T *)scratch;
    la.boxes = la.partial + (size_t)la.J * 2 * a.EN;
    la.live = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(la.boxes + (size_t)la.n_tiles * 5) + 15) & ~uintptr_t(15));
    {   // counters and work list behind the live map, 16-byte aligned
        uintptr_t p = reinterpret_cast<uintptr_t>(la.live) + (size_t)live_bytes;
        la.work_counters = reinterpret_cast<int *>(p);
        la.work = la.work_counters + kWorkCounters;
        la.live_cnt = la.work + ((list_entries + 3) & ~3LL);
        la.live_list = la.live_cnt + ((large_iblocks(a.EN) + 3) & ~3LL);
    }
    if (cur_boxes) la.boxes = (T *)cur_boxes;  // the view's own box table (snp_large_run_p2p)
    // exact culling distance beyond the r+s sums: where exp(rd/B) is identically zero in the arithmetic in use
    const int soc = o->type % 3;
    const double under = sizeof(T) == 8 ? 700.0 : 88.0;
    double margin = -1.0;
    if (!(o->reserved & 2)) {
        if (soc == 0 && c->params[3] > 0) margin = under * c->params[3];
        else if (soc == 1 && c->params[3] > 0 && c->params[7] > 0) margin = under * (c->params[3] > c->params[7] ? c->params[3] : c->params[7]);
    }
    la.cull_margin = (T)margin;
    // culled steps: one agent per thread (twice the units to hand out).  Round 2 kept two for fp32 on the static grid, where the
    // finer split only added CTAs to launch; on the work list one rank's slice of an 8-way split runs 0.0552 -> 0.0408 ms in fp32
    la.apt = margin >= 0.0 ? 1 : kMaxAgentsPerThread;  // fp32 pair evaluations are short: the finer split only adds launch overhead there
    la.use_list = (margin >= 0.0 && self_offset % kTile == 0 && list_entries < (1LL << 31) && !(o->reserved & SNP_OPT_LARGE_GRID)) ? 1 : 0;
    switch (o->type) {
        case 0: return launch_large<T, 0, 0, 0>(la, st);
        case 1: return launch_large<T, 1, 1, 0>(la, st);
        case 2: return launch_large<T, 2, 0, 0>(la, st);
        case 3: return launch_large<T, 0, 0, 1>(la, st);
        case 4: return launch_large<T, 1, 1, 1>(la, st);
        case 5: return launch_large<T, 2, 0, 1>(la, st);
        case 6: return launch_large<T, 0, 0, 2>(la, st);
        case 7: return launch_large<T, 1, 1, 2>(la, st);
        case 8: return launch_large<T, 2, 0, 2>(la, st);
    }
    set_error("Type %d does not exist for this implementation", o->type);
    return SNP_ERR_INVALID;
}

}  // namespace
}  // namespace snp

using namespace snp;

extern "C" {

int64_t snp_large_scratch_bytes(int64_t n_local, int64_t M, int32_t dtype) {
    const long long w = dtype == SNP_F64 ? 8 : 4;
    return w * (large_J(M) * 2 * n_local + large_tiles(M) * 5) + 9 * large_iblocks(n_local) * ((large_J(M) + 15) & ~15LL) + 4 * large_iblocks(n_local) +
           256 + 4 * kWorkCounters;
}

int snp_large_step(const snp_crowd *c, const snp_step_opts *o, const void *others, int64_t M, int64_t self_offset, void *next_view,
                   void *scratch, int64_t scratch_bytes, void *stream) {
    if (!c || !o || !others) { set_error("snp_large_step: null argument"); return SNP_ERR_INVALID; }
    if (!c->dyn || !c->stat || !c->goals || !c->goal_idx || !c->goal_cnt) { set_error("snp_large_step: crowd arrays missing"); return SNP_ERR_INVALID; }
    if (c->agent_params) { set_error("snp_large_step: per-agent parameter rows are not supported"); return SNP_ERR_UNSUPPORTED; }
    if (c->walls_per_env) { set_error("snp_large_step: one wall set per crowd"); return SNP_ERR_INVALID; }
    const long long N = (long long)c->E * c->N;
    if (M < N || self_offset < 0 || self_offset + N > M + 1) { set_error("snp_large_step: M=%lld offset=%lld N=%lld", (long long)M, (long long)self_offset, N); return SNP_ERR_INVALID; }
    if (c->dtype == SNP_F64) return run_large<double>(c, o, others, M, self_offset, next_view, nullptr, 0, scratch, scratch_bytes, (cudaStream_t)stream);
    if (c->dtype == SNP_F32) return run_large<float>(c, o, others, M, self_offset, next_view, nullptr, 0, scratch, scratch_bytes, (cudaStream_t)stream);
    set_error("bad dtype %d", c->dtype);
    return SNP_ERR_INVALID;
}

int snp_large_step_p2p(const snp_crowd *c, const snp_step_opts *o, const void *others, int64_t M, int64_t self_offset,
                       const void *const *peer_next_views, int32_t n_peers, void *scratch, int64_t scratch_bytes, void *stream) {
    if (!c || !o || !others || !peer_next_views) { set_error("snp_large_step_p2p: null argument"); return SNP_ERR_INVALID; }
    if (n_peers < 1 || n_peers > 8) { set_error("snp_large_step_p2p: 1..8 peers (got %d)", n_peers); return SNP_ERR_INVALID; }
    if (!c->dyn || !c->stat || !c->goals || !c->goal_idx || !c->goal_cnt) { set_error("snp_large_step_p2p: crowd arrays missing"); return SNP_ERR_INVALID; }
    if (c->agent_params || c->walls_per_env) { set_error("snp_large_step_p2p: uniform parameters and one wall set only"); return SNP_ERR_UNSUPPORTED; }
    const long long N = (long long)c->E * c->N;
    if (M < N || self_offset < 0 || self_offset + N > M + 1) { set_error("snp_large_step_p2p: M=%lld offset=%lld N=%lld", (long long)M, (long long)self_offset, N); return SNP_ERR_INVALID; }
    if (c->dtype == SNP_F64) return run_large<double>(c, o, others, M, self_offset, nullptr, peer_next_views, n_peers, scratch, scratch_bytes, (cudaStream_t)stream);
    if (c->dtype == SNP_F32) return run_large<float>(c, o, others, M, self_offset, nullptr, peer_next_views, n_peers, scratch, scratch_bytes, (cudaStream_t)stream);
    set_error("bad dtype %d", c->dtype);
    return SNP_ERR_INVALID;
}

int snp_large_run_p2p(const snp_crowd *c, const snp_step_opts *o, const void *const *peer_views_a, const void *const *peer_views_b,
                      int32_t first_is_b, int64_t M, int64_t self_offset, int32_t world, int32_t rank, const void *const *peer_flags,
                      uint64_t epoch_base, int32_t n_substeps, int32_t *error_flag, void *scratch, int64_t scratch_bytes, void *stream) {
    if (!c || !o || !peer_views_a || !peer_views_b || !peer_flags || !error_flag) { set_error("snp_large_run_p2p: null argument"); return SNP_ERR_INVALID; }
    if (world < 1 || world > 8 || rank < 0 || rank >= world) { set_error("snp_large_run_p2p: 1..8 ranks (got world %d rank %d)", world, rank); return SNP_ERR_INVALID; }
    if (!c->dyn || !c->stat || !c->goals || !c->goal_idx || !c->goal_cnt) { set_error("snp_large_run_p2p: crowd arrays missing"); return SNP_ERR_INVALID; }
    if (c->agent_params || c->walls_per_env) { set_error("snp_large_run_p2p: uniform parameters and one wall set only"); return SNP_ERR_UNSUPPORTED; }
    const long long N = (long long)c->E * c->N;
    if (M < N || self_offset < 0 || self_offset + N > M) { set_error("snp_large_run_p2p: M=%lld offset=%lld N=%lld", (long long)M, (long long)self_offset, N); return SNP_ERR_INVALID; }
    if (self_offset % kTile || N % kTile) { set_error("snp_large_run_p2p: every rank's slice must be a whole number of %d-entity tiles", kTile); return SNP_ERR_UNSUPPORTED; }
    if (c->dtype != SNP_F64 && c->dtype != SNP_F32) { set_error("bad dtype %d", c->dtype); return SNP_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t w = c->dtype == SNP_F64 ? 8 : 4;
    // a view buffer = [5][M] entity view followed by its [n_tiles][5] box table
    auto boxes_of = [&](const void *view) { return (const void *)((const char *)view + (size_t)5 * M * w); };
    BarrierArgs b;
    for (int p = 0; p < 8; ++p) b.flags[p] = p < world ? (unsigned long long *)peer_flags[p] : nullptr;
    b.world = world; b.rank = rank; b.error = error_flag;
    int cur_is_b = first_is_b ? 1 : 0;
    for (int s = 0; s < n_substeps; ++s) {
        const void *const *cur = cur_is_b ? peer_views_b : peer_views_a, *const *nxt = cur_is_b ? peer_views_a : peer_views_b;
        const void *nxt_boxes[8];
        for (int p = 0; p < world; ++p) nxt_boxes[p] = boxes_of(nxt[p]);
        int rc;
        // sub-step 0 computes the tile boxes of the view it was handed; from then on the producer has written them
        if (c->dtype == SNP_F64) rc = run_large<double>(c, o, cur[rank], M, self_offset, nullptr, nxt, world, scratch, scratch_bytes, st, boxes_of(cur[rank]), s > 0, nxt_boxes);
        else rc = run_large<float>(c, o, cur[rank], M, self_offset, nullptr, nxt, world, scratch, scratch_bytes, st, boxes_of(cur[rank]), s > 0, nxt_boxes);
        if (rc) return rc;
        if (world > 1) {
            b.epoch = epoch_base + (unsigned long long)s + 1;
            k_rank_barrier<<<1, 32, 0, st>>>(b);
            count_launch();
            SNP_CUDA_OK(cudaGetLastError());
        }
        cur_is_b ^= 1;
    }
    return SNP_OK;
}

int snp_large_publish(const snp_crowd *c, int32_t type, void *view, int64_t stride, int64_t offset, void *stream) {
    if (!c || !view || !c->dyn || !c->stat) { set_error("snp_large_publish: null argument"); return SNP_ERR_INVALID; }
    if (type < 0 || type > 8) { set_error("Type %d does not exist for this implementation", type); return SNP_ERR_INVALID; }
    const long long N = (long long)c->E * c->N;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    if (c->dtype == SNP_F64) k_large_publish<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double *)c->dyn, (const double *)c->stat, N, type >= 3, (double *)view, stride, offset);
    else k_large_publish<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float *)c->dyn, (const float *)c->stat, N, type >= 3, (float *)view, stride, offset);
    count_launch();
    SNP_CUDA_OK(cudaGetLastError());
    return SNP_OK;
}

}  // extern "C"
