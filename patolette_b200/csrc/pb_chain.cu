// pb_chain.cu - ordered ("chain") per-bucket sums.  (The per-cluster mean / covariance /
// distortion passes live in pb_ordered.cu, which adds the binade-speculative fast path.)
//
// Every statistic the reference derives from a cluster - weighted mean
// (array/matrix2D.c:200-233), centred covariance (math/pca.c:84-97), distortion
// (quantize/cluster.c:135-148), per-bucket moments (quantize/local.c:124-134,
// quantize/cells.c:78-116) - is a plain left-to-right f64 accumulation in ascending
// pixel order.  Floating-point addition is not associative, so a tree / warp-shuffle /
// atomic reduction gives different low-order bits (it is MORE accurate, but the
// parity bar is bit-exact, and a flipped low bit can flip a bucket decision three
// stages later).  These kernels therefore keep each accumulation a true sequential
// chain and find their parallelism ACROSS chains:
//   * one warp per cluster; lanes 0..C-1 each own one chain (C = 4 for the mean pass,
//     7 for the centred pass, 4/10 for the per-bucket passes);
//   * all 32 lanes cooperate on the memory side: coalesced (or gathered) loads of a
//     tile into shared memory, double-buffered through registers so the next tile's
//     HBM latency hides behind the current tile's dependent-add chain;
//   * chain lanes then walk the tile reading operands as shared-memory broadcasts -
//     the critical path per element is exactly one dependent DADD.
// Cost model: ~8-10 SM cycles per element per cluster (DADD latency), independent of
// the chain count, so the big early clusters are latency-bound (see DESIGN.md for the
// planned binade-speculative block summaries that lift this to HBM speed while
// staying bit-identical).
#include "pb_common.cuh"
#include "pb_kernels.h"
#include "pb_prof.h"

namespace {

// ------------------------------------------------------------------------------------
// Per-bucket ordered sums.  The members of bucket b of a cluster, in ascending pixel
// order, are ord[class_start[b] .. class_start[b+1]) (stable bucket sort, pb_parallel.cu).
// One warp per (cluster, bucket): lanes gather a tile of members, chain lanes add.
//
// The gathers are what bounds these kernels (measured: 2.2 TB/s of 32-byte sector traffic with the planar
// layout, i.e. the random-access limit of HBM, not the DADD chain).  k_buckets therefore leaves an
// interleaved copy (c0, c1, c2, w) of every pixel behind: one sector per member instead of three.
// The chain lanes run ONE uniform loop `acc += term[lane][e]`: every per-element term is formed by the
// gathering lanes (in parallel) before it is staged, so the critical path per element is a shared-memory
// broadcast read + one dependent DADD.
// ------------------------------------------------------------------------------------
constexpr int BK_WARPS = 4;

struct Px { double c0, c1, c2, w; };
__device__ __forceinline__ Px load_px(const double *__restrict__ aos, uint32_t p) {
    const double2 *a = reinterpret_cast<const double2 *>(aos) + 2 * (size_t)p;
    const double2 u = a[0], v = a[1];
    return Px{u.x, u.y, v.x, v.y};
}

//   LQ (local.c:124-134): term 0 = w (bucket "size"), terms 1..3 = c_j * w.
//     The reference accumulates the size as size_t += double (local.c:133), i.e.
//     size = trunc((double)size + w) at every step.  Unweighted that is the member count (exact), and
//     c_j * 1.0 == c_j: three plain chains.  Weighted, every chain lane runs the trunc-select stream.
//   The gather of tile t+1 sits in registers during the chain over tile t.
constexpr int BS_TILE = 128, BS_STRIDE = BS_TILE + 2, BS_PER = BS_TILE / 32;
template <bool WEIGHTED>
__global__ void __launch_bounds__(BK_WARPS * 32) k_bucket_chains_lq(const double *__restrict__ aos,
                                                                   const PbSeg *__restrict__ segs, int nseg,
                                                                   const uint32_t *__restrict__ ord,
                                                                   const uint32_t *__restrict__ class_start,
                                                                   double *__restrict__ out) {
    constexpr int NT = WEIGHTED ? 4 : 3;
    __shared__ double sm_all[BK_WARPS][NT][BS_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * BK_WARPS + warp;
    if (gw >= nseg * PB_BUCKETS) return;
    const int seg = gw / PB_BUCKETS, b = gw % PB_BUCKETS;
    double(*sm)[BS_STRIDE] = sm_all[warp];
    const uint32_t *cs = class_start + (size_t)seg * (PB_BUCKETS + 1);
    const uint32_t beg = cs[b], end = cs[b + 1];
    double acc = 0.0;
    double g[BS_PER][NT];
    auto gather = [&](uint32_t i0) {
#pragma unroll
        for (int q = 0; q < BS_PER; q++) {
            const uint32_t i = i0 + q * 32 + lane;
#pragma unroll
            for (int t = 0; t < NT; t++) g[q][t] = 0.0;
            if (i < end) {
                const Px x = load_px(aos, ord ? ord[i] : i); // (ord == nullptr: aos is the bucket-sorted copy)
                if (WEIGHTED) {
                    g[q][0] = x.w;
                    g[q][1] = __dmul_rn(x.c0, x.w);
                    g[q][2] = __dmul_rn(x.c1, x.w);
                    g[q][3] = __dmul_rn(x.c2, x.w);
                } else {
                    g[q][0] = x.c0;
                    g[q][1] = x.c1;
                    g[q][2] = x.c2;
                }
            }
        }
    };
    if (beg < end) gather(beg);
    for (uint32_t i0 = beg; i0 < end; i0 += BS_TILE) {
        const uint32_t cnt = min((uint32_t)BS_TILE, end - i0);
#pragma unroll
        for (int q = 0; q < BS_PER; q++)
#pragma unroll
            for (int t = 0; t < NT; t++) sm[t][q * 32 + lane] = g[q][t];
        __syncwarp();
        if (i0 + BS_TILE < end) gather(i0 + BS_TILE); // in flight during the chain below
        if (lane < NT) {
            const double *vp = sm[lane];
            if (WEIGHTED) {
#pragma unroll 8
                for (uint32_t e = 0; e < cnt; e++) {
                    const double t = __dadd_rn(acc, vp[e]);
                    acc = lane == 0 ? trunc(t) : t;
                }
            } else {
#pragma unroll 16
                for (uint32_t e = 0; e < cnt; e++) acc = __dadd_rn(acc, vp[e]);
            }
        }
        __syncwarp();
    }
    double *o = out + ((size_t)seg * PB_BUCKETS + b) * 4;
    if (WEIGHTED) {
        if (lane == 0) o[0] = __longlong_as_double((long long)(unsigned long long)acc);
        else if (lane < 4) o[lane] = acc;
    } else {
        if (lane == 0) o[0] = __longlong_as_double((long long)(unsigned long long)(end - beg));
        if (lane < 3) o[1 + lane] = acc;
    }
}

// GQ cell moments (cells.c:78-116), unweighted, over the whole image: only 512 warps, each with a long
// bucket, so the gathers go through a ring of BK_DEPTH stages filled with cp.async (16-byte copies of the
// interleaved pixels): the gathers of tile t + BK_DEPTH - 1 are issued before the chain over tile t starts,
// and the positions (`ord`) they need were loaded one iteration earlier.
// terms 0..2 = c_j ; 3 = (cx^2 + cy^2) + cz^2 ; 4..9 = c_r * c_s for (r,s) = (0,0)(0,1)(1,1)(0,2)(1,2)(2,2).
constexpr int BK_TILE = 64;            // members per stage (2 per lane)
constexpr int BK_STRIDE = BK_TILE + 2;
constexpr int BK_PER = BK_TILE / 32;
constexpr int BK_DEPTH = 3;            // stages in flight per warp (static shared memory: 45 KB per CTA)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__global__ void __launch_bounds__(BK_WARPS * 32) k_bucket_chains_gq(const double *__restrict__ aos,
                                                                   const uint32_t *__restrict__ ord,
                                                                   const uint32_t *__restrict__ class_start,
                                                                   double *__restrict__ out) {
    constexpr int NT = 10;
    __shared__ __align__(16) double raw_all[BK_WARPS][BK_DEPTH][BK_TILE][4];
    __shared__ double term_all[BK_WARPS][NT][BK_STRIDE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * BK_WARPS + warp;
    if (b >= PB_BUCKETS) return;
    double(*raw)[BK_TILE][4] = raw_all[warp];
    double(*term)[BK_STRIDE] = term_all[warp];
    const uint32_t beg = class_start[b], end = class_start[b + 1];
    double acc = 0.0;
    if (beg < end) {
        const uint32_t ntile = (end - beg + BK_TILE - 1) / BK_TILE;
        auto load_ord = [&](uint32_t t, uint32_t *o) {
#pragma unroll
            for (int q = 0; q < BK_PER; q++) {
                const uint32_t i = beg + t * BK_TILE + q * 32 + lane;
                o[q] = (t < ntile && i < end) ? (ord ? ord[i] : i) : 0xffffffffu;
            }
        };
        auto issue = [&](uint32_t t, const uint32_t *o) { // gathers of tile t into stage t % BK_DEPTH (one commit per call)
#pragma unroll
            for (int q = 0; q < BK_PER; q++) {
                if (o[q] != 0xffffffffu) {
                    const double *src = aos + 4 * (size_t)o[q];
                    double *dst = raw[t % BK_DEPTH][q * 32 + lane];
                    cp_async16(dst, src);
                    cp_async16(dst + 2, src + 2);
                }
            }
            cp_async_commit();
        };
        uint32_t ordv[BK_PER]; // positions of the tile whose gathers are issued next
        {   // prologue: the positions of the first BK_DEPTH tiles in flight together, then their gathers
            uint32_t o[BK_DEPTH][BK_PER];
#pragma unroll
            for (int t = 0; t < BK_DEPTH; t++) load_ord(t, o[t]);
#pragma unroll
            for (int t = 0; t + 1 < BK_DEPTH; t++) issue(t, o[t]);
#pragma unroll
            for (int q = 0; q < BK_PER; q++) ordv[q] = o[BK_DEPTH - 1][q];
        }
        for (uint32_t t = 0; t < ntile; t++) {
            const uint32_t cnt = min((uint32_t)BK_TILE, end - (beg + t * BK_TILE));
            cp_async_wait<BK_DEPTH - 2>(); // tile t has landed (this lane's copies)
            __syncwarp();                  // ... and everybody else's
#pragma unroll
            for (int q = 0; q < BK_PER; q++) {
                const int e = q * 32 + lane;
                const double x = raw[t % BK_DEPTH][e][0], y = raw[t % BK_DEPTH][e][1], z = raw[t % BK_DEPTH][e][2];
                term[0][e] = x;
                term[1][e] = y;
                term[2][e] = z;
                term[3][e] = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
                term[4][e] = __dmul_rn(x, x);
                term[5][e] = __dmul_rn(x, y);
                term[6][e] = __dmul_rn(y, y);
                term[7][e] = __dmul_rn(x, z);
                term[8][e] = __dmul_rn(y, z);
                term[9][e] = __dmul_rn(z, z);
            }
            __syncwarp();
            issue(t + BK_DEPTH - 1, ordv); // refills the stage consumed at iteration t - 1
            load_ord(t + BK_DEPTH, ordv);
            if (lane < NT) {
                const double *vp = term[lane];
#pragma unroll 16
                for (uint32_t e = 0; e < cnt; e++) acc = __dadd_rn(acc, vp[e]);
            }
            __syncwarp();
        }
        cp_async_wait<0>();
    }
    if (lane < 10) out[(size_t)b * 10 + lane] = acc;
}


// GQ cell moments, second version: one CTA per bucket.  With one warp per bucket the gathers of that warp bound
// the kernel (45 cycles per member against 8.4 for the dependent add): here all four warps of the CTA gather -
// two members per thread and tile of 256, the next tile's gathers in registers while the current tile is being
// chained - and warp 0 runs the ten chains (lanes 0..9) over the terms the gatherers staged in shared memory.
constexpr int BG_THREADS = 128, BG_TILE = 256, BG_PER = BG_TILE / BG_THREADS, BG_STRIDE = BG_TILE + 2;
__global__ void __launch_bounds__(BG_THREADS) k_bucket_chains_gq2(const double *__restrict__ aos, const uint32_t *__restrict__ ord,
                                                                 const uint32_t *__restrict__ class_start,
                                                                 double *__restrict__ out) {
    constexpr int NT = 10;
    __shared__ double term[2][NT][BG_STRIDE];
    const int b = blockIdx.x, tid = threadIdx.x;
    const uint32_t beg = class_start[b], end = class_start[b + 1];
    const uint32_t ntile = (end - beg + BG_TILE - 1) / BG_TILE;
    double acc = 0.0;
    Px g[BG_PER];
    auto gather = [&](uint32_t t) {
#pragma unroll
        for (int q = 0; q < BG_PER; q++) {
            const uint32_t i = beg + t * BG_TILE + q * BG_THREADS + tid;
            g[q] = Px{0.0, 0.0, 0.0, 0.0};
            if (t < ntile && i < end) g[q] = load_px(aos, ord ? ord[i] : i);
        }
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int q = 0; q < BG_PER; q++) {
            const int e = q * BG_THREADS + tid;
            const double x = g[q].c0, y = g[q].c1, z = g[q].c2;
            term[buf][0][e] = x;
            term[buf][1][e] = y;
            term[buf][2][e] = z;
            term[buf][3][e] = __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
            term[buf][4][e] = __dmul_rn(x, x);
            term[buf][5][e] = __dmul_rn(x, y);
            term[buf][6][e] = __dmul_rn(y, y);
            term[buf][7][e] = __dmul_rn(x, z);
            term[buf][8][e] = __dmul_rn(y, z);
            term[buf][9][e] = __dmul_rn(z, z);
        }
    };
    gather(0);
    stage(0);
    gather(1);
    __syncthreads();
    for (uint32_t t = 0; t < ntile; t++) {
        // tile t is staged in term[t & 1]; the registers hold tile t + 1: stage it, then fetch tile t + 2 - those
        // loads are in flight while warp 0 chains over tile t
        stage((t + 1) & 1);
        gather(t + 2);
        if (tid < NT) {
            const uint32_t cnt = min((uint32_t)BG_TILE, end - (beg + t * BG_TILE));
            const double *vp = term[t & 1][tid];
#pragma unroll 16
            for (uint32_t e = 0; e < cnt; e++) acc = __dadd_rn(acc, vp[e]);
        }
        __syncthreads();
    }
    if (tid < NT) out[(size_t)b * NT + tid] = acc;
}

} // namespace

static bool g_gq_chain_cta = true; // "gq_chain_cta": one CTA per GQ bucket (k_bucket_chains_gq2) or one warp
void pb_chain_set_gq_cta(bool on) { g_gq_chain_cta = on; }

void pb_launch_bucket_chains_lq(const double *d_aos, const PbSeg *d_segs, int nseg, bool weighted,
                                const uint32_t *d_ord, const uint32_t *d_class_start, double *d_out,
                                cudaStream_t st) {
    if (nseg <= 0) return;
    const int grid = (nseg * PB_BUCKETS + BK_WARPS - 1) / BK_WARPS;
    PbProfScope _prof("k_bucket_chains_lq", st);
    if (weighted) k_bucket_chains_lq<true><<<grid, BK_WARPS * 32, 0, st>>>(d_aos, d_segs, nseg, d_ord, d_class_start, d_out);
    else k_bucket_chains_lq<false><<<grid, BK_WARPS * 32, 0, st>>>(d_aos, d_segs, nseg, d_ord, d_class_start, d_out);
    PB_CUDA_OK(cudaGetLastError());
}

void pb_launch_bucket_chains_gq(const double *d_aos, const uint32_t *d_ord, const uint32_t *d_class_start,
                                double *d_out, cudaStream_t st) {
    { PbProfScope _prof("k_bucket_chains_gq", st);
    if (g_gq_chain_cta) k_bucket_chains_gq2<<<PB_BUCKETS, BG_THREADS, 0, st>>>(d_aos, d_ord, d_class_start, d_out);
    else k_bucket_chains_gq<<<PB_BUCKETS / BK_WARPS, BK_WARPS * 32, 0, st>>>(d_aos, d_ord, d_class_start, d_out);
    }
    PB_CUDA_OK(cudaGetLastError());
}
