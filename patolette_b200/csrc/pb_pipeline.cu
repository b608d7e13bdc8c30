// pb_pipeline.cu - host orchestration of the pixel-array hot path and the C ABI.
//
// Mirrors the stage sequence of the reference's only public entry point
// (lib/src/patolette.c:157-343): colour transform -> GQ (quantize/global.c) -> LQ
// (quantize/local.c) -> palette (palette/create.c | refine.c) -> nearest map
// (palette/nearest.c) | Riemersma dither (dither/riemersma.c) -> palette back to sRGB.
// Every O(N) step is a CUDA kernel (pb_color.cu, pb_chain.cu, pb_parallel.cu,
// pb_kmeans.cu, pb_dither.cu); the host keeps only O(K)/O(512^2) control logic: the
// 3x3 eigen solves, the Wu dynamic programme over 512 buckets and the greedy
// best-first selection.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "../../include/patolette_b200.h"
#include "pb_common.cuh"
#include "pb_host.h"
#include "pb_hostpool.h"
#include "pb_kernels.h"
#include "pb_nccl.h"
#include "pb_dsyev3.h"
#include "pb_pipeline.h"
#include "pb_pool.h"
#include "pb_prof.h"
#include "pb_xfer.h"

namespace {

// The library keeps per-process state (streams, pinned bounce lanes, timings, the buffer cache): calls are
// serialised.  One call saturates the GPU anyway.
std::mutex g_call_mu;
int g_device = 0;
int g_overlap_override = -1; // patolette_b200_set_option "overlap": -1 default, 0 off, 1 on
bool g_nn_grid = true;       // patolette_b200_set_option "nn_grid": candidate-list 1-NN (pb_nngrid.cu) vs brute force
// "split_certify": 1 (default) = the split's optimal bucket comes from unordered per-bucket sums plus a proof that the
// reference's argmax is the same (pb_certify.cu), clusters whose certificate is refused are re-evaluated exactly;
// 0 = every cluster takes the exact route (bucket sort + sequential per-bucket chains); 2 = certified route with every
// certificate refused (tests: exercises the re-evaluation of every cluster)
int g_split_certify = 1;
unsigned long long g_split_redone = 0; // clusters re-evaluated through the exact route since the last reset
bool g_sorted_payload = true; // "sorted_payload": the bucket sort writes the interleaved pixels themselves (sequential per-bucket sums) instead of indices (gathers)

// ---- chain sharding (patolette_b200_set_sharding) --------------------------------------------------
// Every rank holds the whole image and runs the same host logic; the ordered sums - the bulk of the step -
// are split by CHAIN: rank r sums the chains it owns for every segment, and the ranks' PbStats rows are
// all-gathered through the caller's callback (torch.distributed / MPI / anything that can all-gather host
// bytes) and merged field by field.  Every field is still the reference's sequential sum, so the result
// does not depend on the number of ranks.
int g_shard_rank = 0, g_shard_world = 1;
patolette_b200_allgather_fn g_shard_allgather = nullptr;
void *g_shard_user = nullptr;
bool sharded() { return g_shard_world > 1 && g_shard_allgather; }
// chains of the mean pass: {w, c0*w, c1*w, c2*w} (w only when weighted), of the centred pass: 6 covariances + distortion
int mean_owner(int chain, bool weighted) { return (weighted ? chain : chain - 1) % g_shard_world; }
int cent_owner(int chain) { return chain % g_shard_world; }
unsigned mean_mask(bool weighted) {
    if (!sharded()) return ~0u;
    unsigned m = 0;
    for (int c = weighted ? 0 : 1; c < 4; c++)
        if (mean_owner(c, weighted) == g_shard_rank) m |= 1u << c;
    return m;
}
unsigned cent_mask() {
    if (!sharded()) return ~0u;
    unsigned m = 0;
    for (int c = 0; c < 7; c++)
        if (cent_owner(c) == g_shard_rank) m |= 1u << c;
    return m;
}
// all-gather `count` rows and keep, for every field, the owner's value.  Mean pass: the raw sums are scaled
// here (matrix2D.c:230-231: mean = sum * (1 / wsum)) - the same two IEEE operations the kernel performs.
struct pb_shard_error {}; // the caller's all-gather failed: the call ends with exit code -1 on this rank
void shard_merge(PbStats *rows, int count, bool mean_pass, bool weighted) {
    if (!sharded() || count <= 0) return;
    static thread_local std::vector<PbStats> all;
    all.assign((size_t)count * g_shard_world, PbStats{});
    if (g_shard_allgather(rows, all.data(), (size_t)count * sizeof(PbStats), g_shard_user) != 0) throw pb_shard_error{};
    for (int i = 0; i < count; i++) {
        auto from = [&](int rank) -> const PbStats & { return all[(size_t)rank * count + i]; };
        if (mean_pass) {
            if (weighted) rows[i].wsum = from(mean_owner(0, true)).wsum;
            const double inv = 1.0 / rows[i].wsum;
            for (int j = 0; j < 3; j++) rows[i].mean[j] = from(mean_owner(1 + j, weighted)).mean[j] * inv;
        } else {
            for (int t = 0; t < 6; t++) rows[i].cov[t] = from(cent_owner(t)).cov[t];
            rows[i].dist = from(cent_owner(6)).dist;
        }
    }
}
cudaStream_t g_user_stream = nullptr;
bool g_use_user_stream = false;
double g_timings[10] = {0};
double g_saliency_ms = 0.0; // time of the saliency stage of the last call (part of the "color" slot)

template <typename T>
struct DevArr {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = (T *)pb_pool_alloc(count * sizeof(T));
    }
    void release() {
        if (p) pb_pool_free(p);
        p = nullptr;
        n = 0;
    }
    ~DevArr() { release(); }
    DevArr() = default;
    DevArr(const DevArr &) = delete;
    DevArr &operator=(const DevArr &) = delete;
};

struct Timer {
    cudaEvent_t e[2];
    cudaStream_t st;
    explicit Timer(cudaStream_t s) : st(s) {
        cudaEventCreate(&e[0]);
        cudaEventCreate(&e[1]);
    }
    ~Timer() {
        cudaEventDestroy(e[0]);
        cudaEventDestroy(e[1]);
    }
    void start() { cudaEventRecord(e[0], st); }
    double stop() {
        cudaEventRecord(e[1], st);
        cudaEventSynchronize(e[1]);
        float ms = 0;
        cudaEventElapsedTime(&ms, e[0], e[1]);
        return ms;
    }
};

// ------------------------------------------------------------------------------------
// GQ host logic: cumulative cell moments and Wu's dynamic programme
// (quantize/cells.c:118-328, quantize/global.c:72-298).  512 buckets -> O(12 * 512^2).
// ------------------------------------------------------------------------------------
constexpr int CELLS = PB_BUCKETS + 1;
struct CellMoments {
    uint64_t w0[CELLS];
    double w1[3][CELLS];
    double w2[CELLS];
    double wrs[3][3][CELLS]; // [r][s], r <= s
};

inline double sq(double x) { return x * x; }

double cell_distortion(size_t a, size_t b, const CellMoments &m) { // cells.c:141-182
    if (m.w0[a] == m.w0[b]) return 0;
    return m.w2[b] - m.w2[a] -
           (sq(m.w1[0][b] - m.w1[0][a]) + sq(m.w1[1][b] - m.w1[1][a]) + sq(m.w1[2][b] - m.w1[2][a])) /
               (double)(m.w0[b] - m.w0[a]);
}

bool pca_axis_from_vcov(double v[9], double axis[3]) { // pca.c:122-149
    double w[3];
    if (!pb_eigen_solve3(v, w)) return false;
    axis[0] = v[6]; axis[1] = v[7]; axis[2] = v[8];
    return true;
}

bool cell_pca(size_t a, size_t b, const CellMoments &m, double axis[3]) { // cells.c:184-278
    double v[9] = {0};
    for (int s = 0; s < 3; s++)
        for (int r = 0; r <= s; r++) {
            double e = 0;
            if (m.w0[a] != m.w0[b]) {
                double cnt = (double)(m.w0[b] - m.w0[a]);
                e = (m.wrs[r][s][b] - m.wrs[r][s][a]) / cnt -
                    (m.w1[r][b] - m.w1[r][a]) * (m.w1[s][b] - m.w1[s][a]) / sq(cnt);
            }
            v[s * 3 + r] = e;
        }
    v[0 * 3 + 2] = v[2 * 3 + 0];
    v[0 * 3 + 1] = v[1 * 3 + 0];
    v[1 * 3 + 2] = v[2 * 3 + 1];
    return pca_axis_from_vcov(v, axis);
}

double norm3(const double v[3]) { // vector.c:135-159
    double s = 0;
    for (int i = 0; i < 3; i++) s += pow(v[i], 2);
    return sqrt(s);
}

double cell_bias(size_t a, size_t b, const double axis[3], const CellMoments &m) { // cells.c:280-328
    double ca[3];
    if (!cell_pca(a, b, m, ca)) return -1;
    double norms = norm3(axis) * norm3(ca);
    if (norms < PB_DELTA) return 0;
    double dot = ca[0] * axis[0] + ca[1] * axis[1] + ca[2] * axis[2];
    return fmin(1, fabs(dot / norms));
}

bool gq_should_terminate(const size_t *q, size_t qlen, const double axis[3], const CellMoments &m) {
    // global.c:99-187 (thresholds :19-21)
    double distortion = 0;
    for (size_t j = 0; j + 1 < qlen; j++) distortion += cell_distortion(q[j], q[j + 1], m);
    if (distortion < PB_DELTA) return true;
    double bias = 0;
    for (size_t i = 0; i + 1 < qlen; i++) {
        double cd = cell_distortion(q[i], q[i + 1], m);
        double cb = cell_bias(q[i], q[i + 1], axis, m);
        if (cb < 0) return true;
        if (cb < 0.9) continue;
        bias += (cd / distortion) * cb;
    }
    return bias < 0.1;
}

void l_chain(const std::vector<double> &L, size_t ld, size_t k, size_t N, size_t *chain) { // global.c:72-97
    size_t t = N;
    for (size_t j = k - 1; j >= 1; j--) {
        t = (size_t)L[t * ld + (j + 1)];
        chain[j] = t;
    }
    chain[0] = 0;
    chain[k] = N;
}

// cells.c:53-139: bucket b lives in 1-based slot b + 1; then prefix sums.  hs: 512 x {sum c[3], sum |c|^2,
// sum c_r*c_s for (r,s) = (0,0)(0,1)(1,1)(0,2)(1,2)(2,2)}, hcs: 513 class starts.
void build_cell_moments(const double *hs, const uint32_t *hcs, CellMoments &m) {
    memset(&m, 0, sizeof m);
    for (int b = 0; b < PB_BUCKETS; b++) {
        const double *s = &hs[(size_t)b * 10];
        m.w0[b + 1] = hcs[b + 1] - hcs[b];
        m.w1[0][b + 1] = s[0]; m.w1[1][b + 1] = s[1]; m.w1[2][b + 1] = s[2];
        m.w2[b + 1] = s[3];
        m.wrs[0][0][b + 1] = s[4]; m.wrs[0][1][b + 1] = s[5]; m.wrs[1][1][b + 1] = s[6];
        m.wrs[0][2][b + 1] = s[7]; m.wrs[1][2][b + 1] = s[8]; m.wrs[2][2][b + 1] = s[9];
    }
    for (int i = 1; i < CELLS; i++) {
        m.w0[i] += m.w0[i - 1];
        m.w2[i] += m.w2[i - 1];
        for (int j = 0; j < 3; j++) m.w1[j][i] += m.w1[j][i - 1];
        for (int s = 0; s < 3; s++)
            for (int r = 0; r <= s; r++) m.wrs[r][s][i] += m.wrs[r][s][i - 1];
    }
}

int g_gq_threads = 0; // 0: pb_hostpool_default_threads(); patolette_b200_set_option "gq_threads"
bool g_gq_full_table = false; // "gq_full_table": the reference's (max(K, 512) + 1)^2 layout of L instead of 513 x 13

// Wu's dynamic programme over the 512 buckets (global.c:189-298).  E_k[n] = min_t E_{k-1}[t] + D(t, n) is
// independent for every n, so the n loop of an iteration is spread over a few host threads (each n keeps
// its own descending scan over t, i.e. exactly the reference's comparisons and tie-breaks): 8.8 ms -> ~1.5 ms
// for the full 11 iterations an image with structure goes through.  The reference's L table is
// (max(K, 512) + 1)^2 doubles of which only the first 13 columns are ever read; here it has those.
size_t principal_quantizer(size_t K, const CellMoments &m, size_t *q) {
    const size_t N = CELLS - 1, max_k = 12;
    double axis[3];
    if (!cell_pca(0, N, m, axis)) return 0;
    std::vector<double> E(N + 1, 0.0), E2(N + 1, 0.0);
    const size_t ls = g_gq_full_table ? std::max(K, N) + 1 : max_k + 1; // (the reference's layout, for the self-test)
    std::vector<double> L;
    try {
        L.assign((g_gq_full_table ? ls : N + 1) * ls, 0.0);
    } catch (...) { return 0; }
    for (size_t i = 1; i <= N; i++) E[i] = cell_distortion(0, i, m);
    for (size_t i = 1; i <= K && i < ls; i++) L[i * ls + i] = (double)i;
    size_t k_out = 1;
    l_chain(L, ls, 1, N, q);
    const size_t kmax = std::min(max_k, K);
    // (ranks of an image-sharded job share the host's cores: each takes its share)
    const int threads = g_gq_threads > 0 ? g_gq_threads : std::max(1, pb_hostpool_default_threads() / std::max(1, pb_nccl_world()));
    for (size_t k = 2; k <= kmax; k++) {
        if (gq_should_terminate(q, k_out + 1, axis, m)) break;
        E2 = E;
        pb_hostpool_run(threads, [&](int tid) {
            for (size_t n = k + 1 + (size_t)tid; n <= N; n += (size_t)threads) { // interleaved: the work grows with n
                double cut = (double)(n - 1);
                double e = E2[n - 1];
                for (size_t t = n - 2; t >= k - 1; t--) {
                    double c = E2[t] + cell_distortion(t, n, m);
                    if (c < e) { cut = (double)t; e = c; }
                }
                L[n * ls + k] = cut;
                E[n] = e;
            }
        });
        l_chain(L, ls, k, N, q);
        k_out = k;
    }
    return k_out;
}

// ------------------------------------------------------------------------------------
// Device workspace for one image.
// ------------------------------------------------------------------------------------
struct HNode {
    PbSeg seg;
    PbStats st;
    int owner = -1; // image-sharded runs: the rank that holds this cluster's pixels, -1 = every rank does
};

// ---- image-sharded runs (patolette_b200_sharded, NCCL) ---------------------------------------------------
// Rank r brings pixels [first, first + count) of the image.  After the colour transform the planes are
// all-gathered, GQ runs replicated, and the split loop is sharded by CLUSTER: split_cluster(c) is a pure
// function of c's pixels, so which GPU evaluates a cluster cannot change a bit of the result.  While the tree is
// narrower than the machine every rank evaluates every cluster (the data stays replicated); from the first batch
// with at least `world` replicated clusters on, clusters are dealt out (largest first, to the least loaded rank)
// and a cluster's children stay with the rank that made them - that rank is the only one holding their pixels.
// Per batch the ranks all-gather the children's segments and statistics (240 B per evaluated cluster) on the
// device; the greedy selection runs replicated on every host, on identical numbers.
struct ShardCtx {
    bool on = false;
    int rank = 0, world = 1;
    size_t S = 0;     // slice stride: ceil(n / world) rounded up to 1024 pixels
    size_t first = 0; // first pixel of this rank
    size_t count = 0; // pixels of this rank
};
size_t shard_stride(size_t n, int world) {
    const size_t per = (n + (size_t)world - 1) / (size_t)world;
    return (per + 1023) & ~(size_t)1023;
}
ShardCtx make_shard_ctx(size_t n, bool on) {
    ShardCtx c;
    if (!on || !pb_nccl_active()) return c;
    c.on = true;
    c.rank = pb_nccl_rank();
    c.world = pb_nccl_world();
    c.S = shard_stride(n, c.world);
    c.first = std::min(n, (size_t)c.rank * c.S);
    c.count = std::min(c.S, n - c.first);
    return c;
}
struct XRes { // what the evaluation of one cluster produces
    PbSeg seg[2];
    PbStats st[2];
};
__global__ void k_pack_results(const PbSeg *__restrict__ children, const PbStats *__restrict__ stats,
                               const int *__restrict__ gidx, int nb, XRes *__restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    XRes x;
    x.seg[0] = children[2 * b]; x.seg[1] = children[2 * b + 1];
    x.st[0] = stats[2 * b]; x.st[1] = stats[2 * b + 1];
    out[gidx[b]] = x;
}
struct HPair {
    bool valid = false;
    HNode l, r;
};

struct Quantizer {
    size_t N = 0;
    bool weighted = false;
    int sm_count = 148;
    cudaStream_t st = nullptr;
    bool own_stream = true;
    long launches = 0;

    DevArr<double> col[3], wgt;                // original order, quantisation space
    DevArr<double> pc[2][3], pw[2];            // ping-pong permuted planes
    DevArr<uint32_t> pidx[2];
    DevArr<uint16_t> bucket;
    DevArr<double> aos; // interleaved (c0, c1, c2, w) copy written by k_buckets for the gathering bucket sums
    DevArr<uint32_t> ord, tile_hist, cstart_b, cstart_s;
    DevArr<double> bsums, axes;
    DevArr<PbSeg> segs, children;
    DevArr<PbStats> stats;
    DevArr<PbSplit> split;
    DevArr<PbHist> hist; // certified route: per-cluster bucket tables
    DevArr<uint8_t> lut;
    DevArr<char> oscratch; // block sums + summaries of the speculative ordered sums
    PbPlanes orig{}, bufs[2]{};
    ShardCtx sh;                       // image-sharded run (NCCL) or not
    DevArr<XRes> xres, xall;           // this rank's batch results by global batch position; all ranks' (all-gathered)
    DevArr<int> gidx[2];               // per half-batch: global batch position of every local cluster

    // Second half-batch context.  The ordered-sum resolve of a batch is latency-bound (one warp per chain,
    // a few dozen busy SMs) while the summaries, scatters and bucket chains are throughput work: a batch is
    // therefore split into two halves on two streams, so that one half's resolve overlaps the other half's
    // streaming kernels.  Half 0 uses the members above and `st`.
    struct Half {
        DevArr<uint32_t> tile_hist, cstart_b, cstart_s;
        DevArr<double> bsums, axes;
        DevArr<PbSeg> segs, children;
        DevArr<PbStats> stats;
        DevArr<PbSplit> split;
        DevArr<PbHist> hist;
        DevArr<char> oscratch;
        cudaStream_t st = nullptr;
        cudaEvent_t fork = nullptr;
    } hx;
    bool overlap = true;

#ifndef PB_MAXB
#define PB_MAXB 64
#endif
    static constexpr int MAXB = PB_MAXB; // clusters evaluated per batch (their 2 * MAXB children get stats)
    size_t max_blocks = 0;         // capacity of the packed ordered-sum block table
    size_t max_tiles = 0;          // capacity of the packed scatter tile table

    void init(size_t n, bool with_weights, size_t n_alloc = 0) {
        if (n_alloc < n) n_alloc = n;
        N = n;
        weighted = with_weights;
        PB_CUDA_OK(cudaSetDevice(g_device));
        static int cached_dev = -1, cached_sms = 0;
        if (cached_dev != g_device) {
            PB_CUDA_OK(cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, g_device));
            cached_dev = g_device;
        }
        sm_count = cached_sms;
        if (g_use_user_stream) { st = g_user_stream; own_stream = false; }
        else PB_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (int j = 0; j < 3; j++) col[j].alloc(n_alloc);
        if (weighted) wgt.alloc(n_alloc);
        orig = PbPlanes{{col[0].p, col[1].p, col[2].p}, nullptr, nullptr};
    }
    void init_tree() {
        for (int b = 0; b < 2; b++) {
            for (int j = 0; j < 3; j++) pc[b][j].alloc(N);
            if (weighted) pw[b].alloc(N);
            pidx[b].alloc(N);
            bufs[b] = PbPlanes{{pc[b][0].p, pc[b][1].p, pc[b][2].p}, weighted ? pw[b].p : nullptr, pidx[b].p};
        }
        bucket.alloc(N);
        aos.alloc(4 * N);
        ord.alloc(N);
        max_tiles = pb_scatter_tiles((uint32_t)N) + MAXB;
        tile_hist.alloc(pb_scatter_table_words(max_tiles, MAXB, PB_BUCKETS));
        cstart_b.alloc(MAXB * (PB_BUCKETS + 1));
        cstart_s.alloc(MAXB * 17);
        bsums.alloc(MAXB * PB_BUCKETS * 10);
        axes.alloc(MAXB * 3);
        segs.alloc(MAXB);
        children.alloc(2 * MAXB);
        stats.alloc(2 * MAXB);
        split.alloc(MAXB);
        hist.alloc(MAXB);
        lut.alloc(PB_BUCKETS);
        if (sh.on) {
            xres.alloc(MAXB);
            xall.alloc((size_t)MAXB * sh.world);
            gidx[0].alloc(MAXB);
            gidx[1].alloc(MAXB);
        }
        max_blocks = 2 * ((size_t)pb_ordered_blocks((uint32_t)N) + 2 * MAXB) + 64;
        oscratch.alloc(pb_ordered_scratch_bytes(max_blocks));
        const char *e = getenv("PB200_OVERLAP");
        overlap = !(e && e[0] == '0') && N >= (size_t)1 << 16;
        if (g_overlap_override >= 0) overlap = g_overlap_override != 0;
        if (overlap) {
            hx.tile_hist.alloc(pb_scatter_table_words(max_tiles, MAXB, PB_BUCKETS));
            hx.cstart_b.alloc(MAXB * (PB_BUCKETS + 1));
            hx.cstart_s.alloc(MAXB * 17);
            hx.bsums.alloc(MAXB * PB_BUCKETS * 10);
            hx.axes.alloc(MAXB * 3);
            hx.segs.alloc(MAXB);
            hx.children.alloc(2 * MAXB);
            hx.stats.alloc(2 * MAXB);
            hx.split.alloc(MAXB);
            hx.hist.alloc(MAXB);
            hx.oscratch.alloc(pb_ordered_scratch_bytes(max_blocks));
            // higher priority than the first half's stream: its CTAs are placed first, which staggers the two
            // halves (one streams at full speed while the other is inside its latency-bound resolve)
            int prio_least = 0, prio_greatest = 0;
            PB_CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
            PB_CUDA_OK(cudaStreamCreateWithPriority(&hx.st, cudaStreamNonBlocking, prio_greatest));
            PB_CUDA_OK(cudaEventCreateWithFlags(&hx.fork, cudaEventDisableTiming));
        }
    }
    ~Quantizer() {
        if (hx.st) cudaStreamDestroy(hx.st);
        if (hx.fork) cudaEventDestroy(hx.fork);
        if (st && own_stream) cudaStreamDestroy(st);
    }
    void sync() { PB_CUDA_OK(cudaStreamSynchronize(st)); }
    template <typename T>
    void h2d(T *dst, const T *src, size_t count) {
        PB_CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    }
    template <typename T>
    void d2h(T *dst, const T *src, size_t count) {
        PB_CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, st));
    }

    // ---- GQ (global.c:388-443): returns cluster count, 0 on error -------------------
    size_t run_gq(size_t K, std::vector<HNode> &out) {
        PbSeg whole{0u, (uint32_t)N, 0u, 0u, 0u, 0u};
        PbStats hst;
        const PbPlanes gq[2] = {orig, orig};
        h2d(segs.p, &whole, 1);
        pb_prof_next_bytes(24.0 * N);
        pb_launch_pass_mean(gq, segs.p, 1, (uint32_t)N, (uint32_t)max_blocks, false, stats.p, oscratch.p, oscratch.n, st,
                            mean_mask(false), sharded());     // global.c:407: UNWEIGHTED PCA
        if (sharded()) { // the centred pass needs the complete mean on every rank
            d2h(&hst, stats.p, 1);
            sync();
            shard_merge(&hst, 1, true, false);
            h2d(stats.p, &hst, 1);
        }
        pb_prof_next_bytes(24.0 * N);
        pb_launch_pass_centered(gq, segs.p, 1, (uint32_t)N, (uint32_t)max_blocks, false, stats.p, oscratch.p, oscratch.n, st,
                                cent_mask());
        d2h(&hst, stats.p, 1);
        sync();
        shard_merge(&hst, 1, false, false);
        double v[9], axis[3];
        fill_vcov(hst, v);
        if (!pca_axis_from_vcov(v, axis)) return 0;
        h2d(axes.p, axis, 3);
        pb_prof_next_bytes(24.0 * N);
        pb_launch_dots_minmax(gq, segs.p, 1, (uint32_t)N, axes.p, split.p, sm_count, st);
        pb_prof_next_bytes(26.0 * N);
        pb_launch_buckets(gq, segs.p, 1, (uint32_t)N, axes.p, split.p, bucket.p, g_sorted_payload ? nullptr : aos.p, sm_count, st);
        pb_launch_class_rank(PB_CLS_BUCKET, PB_BUCKETS, segs.p, 1, (uint32_t)N, max_tiles, bucket.p, split.p, lut.p,
                             tile_hist.p, cstart_b.p, st);
        if (g_sorted_payload) {
            pb_prof_next_bytes((24.0 + 2.0 + 32.0) * N);
            pb_launch_scatter_sorted(PB_CLS_BUCKET, PB_BUCKETS, gq, segs.p, 1, (uint32_t)N, bucket.p, split.p, lut.p, tile_hist.p,
                                     cstart_b.p, aos.p, st);
        } else
            pb_launch_scatter_ord(PB_CLS_BUCKET, PB_BUCKETS, segs.p, 1, (uint32_t)N, bucket.p, split.p, lut.p,
                                  tile_hist.p, cstart_b.p, ord.p, st);
        pb_launch_bucket_chains_gq(aos.p, g_sorted_payload ? nullptr : ord.p, cstart_b.p, bsums.p, st);
        std::vector<double> hs(PB_BUCKETS * 10);
        std::vector<uint32_t> hcs(PB_BUCKETS + 1);
        d2h(hs.data(), bsums.p, hs.size());
        d2h(hcs.data(), cstart_b.p, hcs.size());
        sync();
        static thread_local CellMoments m;
        build_cell_moments(hs.data(), hcs.data(), m);
        size_t q[16];
        size_t cells = principal_quantizer(K, m, q);
        if (cells == 0) return 0;
        // global.c:322-332: bucket -> first cell j with bucket + 1 <= q[j + 1]
        uint8_t hl[PB_BUCKETS];
        for (size_t b = 0; b < PB_BUCKETS; b++) {
            hl[b] = 0;
            for (size_t j = 0; j < cells; j++)
                if (b + 1 <= q[j + 1]) { hl[b] = (uint8_t)j; break; }
        }
        h2d(lut.p, hl, PB_BUCKETS);
        // global.c:334-356: per-cell ascending index lists == stable scatter by cell
        pb_launch_class_rank(PB_CLS_LUT, (int)cells, segs.p, 1, (uint32_t)N, max_tiles, bucket.p, split.p, lut.p,
                             tile_hist.p, cstart_s.p, st);
        PbPlanes src = orig;
        src.w = weighted ? wgt.p : nullptr;
        const PbPlanes srcs[2] = {src, src};
        pb_launch_scatter_payload(PB_CLS_LUT, (int)cells, srcs, bufs, true, segs.p, 1, (uint32_t)N, bucket.p,
                                  split.p, lut.p, tile_hist.p, cstart_s.p, st);
        std::vector<uint32_t> cst(cells + 1);
        d2h(cst.data(), cstart_s.p, cells + 1);
        sync();
        out.resize(cells);
        std::vector<PbSeg> hsegs(cells);
        uint32_t bb = 0;
        for (size_t j = 0; j < cells; j++) {
            hsegs[j] = PbSeg{cst[j], cst[j + 1] - cst[j], 0u, 0u, bb, 0u};
            bb += pb_ordered_blocks(hsegs[j].n) + 1;
        }
        h2d(segs.p, hsegs.data(), cells);
        uint32_t cmax = 0;
        for (size_t j = 0; j < cells; j++) cmax = std::max(cmax, hsegs[j].n);
        pb_prof_next_bytes((weighted ? 32.0 : 24.0) * N);
        std::vector<PbStats> hstats(cells);
        pb_launch_pass_mean(bufs, segs.p, (int)cells, cmax, (uint32_t)max_blocks, weighted, stats.p, oscratch.p, oscratch.n, st,
                            mean_mask(weighted), sharded());
        if (sharded()) {
            d2h(hstats.data(), stats.p, cells);
            sync();
            shard_merge(hstats.data(), (int)cells, true, weighted);
            h2d(stats.p, hstats.data(), cells);
        }
        pb_prof_next_bytes((weighted ? 32.0 : 24.0) * N);
        pb_launch_pass_centered(bufs, segs.p, (int)cells, cmax, (uint32_t)max_blocks, weighted, stats.p, oscratch.p, oscratch.n, st,
                                cent_mask());
        d2h(hstats.data(), stats.p, cells);
        sync();
        shard_merge(hstats.data(), (int)cells, false, weighted);
        for (size_t j = 0; j < cells; j++) out[j] = HNode{hsegs[j], hstats[j]};
        return cells;
    }

    static void fill_vcov(const PbStats &s, double v[9]) {
        // pca.c:84-97: V(j,k) = sum / w_sum; column-major, dsyev reads the lower triangle (j >= k)
        static const int jj[6] = {0, 1, 1, 2, 2, 2}, kk[6] = {0, 0, 1, 0, 1, 2};
        for (int t = 0; t < 6; t++) {
            double e = s.cov[t] / s.wsum;
            v[kk[t] * 3 + jj[t]] = e;
            v[jj[t] * 3 + kk[t]] = e;
        }
    }

    // ---- split_cluster (local.c:179-254) for a batch of clusters ------------------------
    // For every node: principal axis (host 3x3 eigen solve), projections, 512 buckets, ordered
    // per-bucket sums, optimal cut, stable partition into the other ping-pong buffer, and the
    // mean / covariance / distortion of both children.  One host synchronisation per batch.
    void eval_split(HNode *const nodes[], HPair *const outs[], int count) {
        // (chain-sharded runs keep the exact route: every rank must make the same sequence of exchanges)
        const bool certify = g_split_certify != 0 && !sharded();
        eval_split_impl(nodes, outs, count, !certify);
        if (!certify) return;
        // clusters whose certificate was refused: once more, through the exact route.  The parent's planes are
        // intact (a partition writes the OTHER buffer), the children's ranges and scratch are simply rewritten.
        HNode *rn[MAXB];
        HPair *ro[MAXB];
        int rc = 0;
        for (int i = 0; i < count; i++)
            if (outs[i]->valid && outs[i]->l.seg.pad == PB_ROUTE_REFUSED) { rn[rc] = nodes[i]; ro[rc++] = outs[i]; }
        if (rc) {
            g_split_redone += (unsigned long long)rc;
            eval_split_impl(rn, ro, rc, true);
        }
    }
    void eval_split_impl(HNode *const nodes[], HPair *const outs[], int count, bool exact) {
        struct Cand { int i; double axis[3]; };
        Cand cand[MAXB];
        int nc = 0;
        for (int i = 0; i < count; i++) {
            outs[i]->valid = false;
            if (nodes[i]->seg.n <= 1) continue; // local.c:187
            double v[9];
            fill_vcov(nodes[i]->st, v);
            if (!pca_axis_from_vcov(v, cand[nc].axis)) continue; // local.c:193-196
            cand[nc++].i = i;
        }
        if (nc == 0) return;
        std::sort(cand, cand + nc, [&](const Cand &a, const Cand &b) { return nodes[a.i]->seg.n > nodes[b.i]->seg.n; });
        // image-sharded runs: the rank that evaluates every candidate (-1: all of them, the result stays replicated).
        // Every rank runs this on identical inputs, so all agree.  Replicated clusters are dealt out as soon as there
        // are enough of them for every rank (or the batch mixes them with clusters that already have an owner).
        int exec[MAXB];
        for (int c = 0; c < nc; c++) exec[c] = -1;
        if (sh.on) {
            double load[256] = {0};
            int n_rep = 0;
            bool any_owned = false;
            for (int c = 0; c < nc; c++) {
                const HNode &nd = *nodes[cand[c].i];
                if (nd.owner >= 0) { exec[c] = nd.owner; load[nd.owner] += nd.seg.n; any_owned = true; }
                else n_rep++;
            }
            if (n_rep >= sh.world || any_owned)
                for (int c = 0; c < nc; c++) { // largest first (cand is sorted), each to the least loaded rank
                    if (exec[c] >= 0) continue;
                    int best = 0;
                    for (int r = 1; r < sh.world; r++)
                        if (load[r] < load[best]) best = r;
                    exec[c] = best;
                    load[best] += nodes[cand[c].i]->seg.n;
                }
        }
        auto mine = [&](int c) { return exec[c] < 0 || exec[c] == sh.rank; };
        int nlocal = 0;
        for (int c = 0; c < nc; c++) nlocal += mine(c) ? 1 : 0;
        // two half-batches (largest first, each to the lighter half); one when profiling per kernel
        const int nhalf = (overlap && nlocal >= 2 && (!pb_prof_enabled() || pb_prof_timeline())) ? 2 : 1;
        struct HalfBatch {
            PbSeg hsegs[MAXB];
            double haxes[3 * MAXB];
            int gidx[MAXB]; // position of the cluster in the (rank-independent) sorted candidate list
            int map[MAXB], nb = 0;
            uint32_t max_n = 0, tb = 0, bb = 0;
            double tot_n = 0;
            PbSeg hch[2 * MAXB];
            PbStats hst[2 * MAXB];
        };
        static thread_local HalfBatch hb[2];
        for (int h = 0; h < 2; h++) { hb[h].nb = 0; hb[h].max_n = hb[h].tb = hb[h].bb = 0; hb[h].tot_n = 0; }
        for (int c = 0; c < nc; c++) {
            if (!mine(c)) continue;
            HalfBatch &B = hb[(nhalf == 2 && hb[1].tot_n < hb[0].tot_n) ? 1 : 0];
            const HNode &nd = *nodes[cand[c].i];
            PbSeg &sg = B.hsegs[B.nb];
            sg = nd.seg;
            sg.tbase = B.tb;
            sg.bbase = B.bb;
            // weighted certified route: every partial bucket size stays below sum(w) < 2^E
            sg.pad = (weighted && !exact) ? (uint32_t)std::min(64, std::max(0, ilogb(nd.st.wsum) + 2)) : 0u;
            B.tb += (uint32_t)pb_scatter_tiles(sg.n);
            B.bb += pb_ordered_blocks(sg.n) + 1;
            memcpy(&B.haxes[3 * B.nb], cand[c].axis, sizeof cand[c].axis);
            B.max_n = std::max(B.max_n, sg.n);
            B.tot_n += sg.n;
            B.gidx[B.nb] = c;
            B.map[B.nb++] = cand[c].i;
        }
        struct Dev { // device scratch of a half
            PbSeg *segs, *children; double *axes, *bsums; PbSplit *split; PbStats *stats;
            uint32_t *tile_hist, *cstart_b, *cstart_s; char *oscratch; size_t oscratch_n; cudaStream_t st; PbHist *hist;
        };
        const Dev dev[2] = {
            {segs.p, children.p, axes.p, bsums.p, split.p, stats.p, tile_hist.p, cstart_b.p, cstart_s.p, oscratch.p, oscratch.n, st, hist.p},
            {hx.segs.p, hx.children.p, hx.axes.p, hx.bsums.p, hx.split.p, hx.stats.p, hx.tile_hist.p, hx.cstart_b.p, hx.cstart_s.p,
             hx.oscratch.p, hx.oscratch.n, hx.st, hx.hist.p}};
        if (nhalf == 2) { // the second stream starts after everything already queued on the first
            PB_CUDA_OK(cudaEventRecord(hx.fork, st));
            PB_CUDA_OK(cudaStreamWaitEvent(hx.st, hx.fork, 0));
        }
        const double bpp = weighted ? 32.0 : 24.0; // planar f64 payload per pixel
        const PbPlanes swapped[2] = {bufs[1], bufs[0]};
        // the launch sequence of split_cluster, step by step and alternating between the halves
        auto step = [&](int k, const HalfBatch &B, const Dev &d) {
            const int nb = B.nb;
            const uint32_t max_n = B.max_n;
            switch (k) {
            case 0:
                PB_CUDA_OK(cudaMemcpyAsync(d.segs, B.hsegs, nb * sizeof(PbSeg), cudaMemcpyHostToDevice, d.st));
                PB_CUDA_OK(cudaMemcpyAsync(d.axes, B.haxes, 3 * nb * sizeof(double), cudaMemcpyHostToDevice, d.st));
                pb_prof_next_bytes(24.0 * B.tot_n);
                pb_launch_dots_minmax(bufs, d.segs, nb, max_n, d.axes, d.split, sm_count, d.st);
                if (!exact) { // certified route: bucket ids + unordered per-bucket sums, no sort
                    pb_prof_next_bytes((bpp + 2.0) * B.tot_n);
                    pb_launch_buckets_hist(bufs, d.segs, nb, max_n, weighted, d.axes, d.split, bucket.p, d.hist, sm_count, d.st);
                    break;
                }
                pb_prof_next_bytes((26.0 + (g_sorted_payload ? 0.0 : 32.0)) * B.tot_n);
                pb_launch_buckets(bufs, d.segs, nb, max_n, d.axes, d.split, bucket.p, g_sorted_payload ? nullptr : aos.p, sm_count, d.st);
                break;
            case 1:
                if (!exact) break;
                pb_launch_class_rank(PB_CLS_BUCKET, PB_BUCKETS, d.segs, nb, max_n, max_tiles, bucket.p, d.split, lut.p,
                                     d.tile_hist, d.cstart_b, d.st);
                if (g_sorted_payload) {
                    pb_prof_next_bytes((bpp + 2.0 + 32.0) * B.tot_n);
                    pb_launch_scatter_sorted(PB_CLS_BUCKET, PB_BUCKETS, bufs, d.segs, nb, max_n, bucket.p, d.split, lut.p, d.tile_hist,
                                             d.cstart_b, aos.p, d.st);
                } else
                    pb_launch_scatter_ord(PB_CLS_BUCKET, PB_BUCKETS, d.segs, nb, max_n, bucket.p, d.split, lut.p, d.tile_hist,
                                          d.cstart_b, ord.p, d.st);
                break;
            case 2:
                if (!exact) {
                    pb_launch_split_certify(d.hist, nb, weighted, d.split, g_split_certify == 2 ? 1 : 0, d.st);
                    break;
                }
                pb_prof_next_bytes(32.0 * B.tot_n);
                pb_launch_bucket_chains_lq(aos.p, d.segs, nb, weighted, g_sorted_payload ? nullptr : ord.p, d.cstart_b, d.bsums, d.st);
                pb_launch_split_select(d.bsums, d.cstart_b, nb, d.split, d.st);
                break;
            case 3:
                pb_launch_class_rank(PB_CLS_SPLIT, 2, d.segs, nb, max_n, max_tiles, bucket.p, d.split, lut.p, d.tile_hist,
                                     d.cstart_s, d.st);
                pb_prof_next_bytes((2 * (bpp + 4.0) + 2.0) * B.tot_n);
                pb_launch_scatter_payload(PB_CLS_SPLIT, 2, bufs, swapped, false, d.segs, nb, max_n, bucket.p, d.split,
                                          lut.p, d.tile_hist, d.cstart_s, d.st);
                pb_launch_make_children(d.segs, nb, d.split, d.children, d.st);
                break;
            case 4:
                pb_prof_next_bytes(bpp * B.tot_n);
                pb_launch_pass_mean(bufs, d.children, 2 * nb, max_n, (uint32_t)max_blocks, weighted, d.stats, d.oscratch,
                                    d.oscratch_n, d.st, mean_mask(weighted), sharded());
                break;
            case 5:
                pb_prof_next_bytes(bpp * B.tot_n);
                pb_launch_pass_centered(bufs, d.children, 2 * nb, max_n, (uint32_t)max_blocks, weighted, d.stats,
                                        d.oscratch, d.oscratch_n, d.st, cent_mask());
                break;
            }
        };
        for (int k = 0; k < 6; k++) {
            if (k == 5 && sharded()) { // between the mean and the centred pass: complete means on every rank
                for (int h = 0; h < nhalf; h++)
                    if (hb[h].nb)
                        PB_CUDA_OK(cudaMemcpyAsync(hb[h].hst, dev[h].stats, 2 * hb[h].nb * sizeof(PbStats), cudaMemcpyDeviceToHost, dev[h].st));
                for (int h = 0; h < nhalf; h++) {
                    if (!hb[h].nb) continue;
                    PB_CUDA_OK(cudaStreamSynchronize(dev[h].st));
                    shard_merge(hb[h].hst, 2 * hb[h].nb, true, weighted);
                    PB_CUDA_OK(cudaMemcpyAsync(dev[h].stats, hb[h].hst, 2 * hb[h].nb * sizeof(PbStats), cudaMemcpyHostToDevice, dev[h].st));
                }
            }
            for (int h = 0; h < nhalf; h++)
                if (hb[h].nb) step(k, hb[h], dev[h]);
        }
        if (sh.on) {
            // every rank packs what it evaluated into the batch-position slots, the slots are all-gathered on the
            // device, and one copy brings every rank's results home
            for (int h = 0; h < nhalf; h++) {
                if (!hb[h].nb) continue;
                PB_CUDA_OK(cudaMemcpyAsync(gidx[h].p, hb[h].gidx, hb[h].nb * sizeof(int), cudaMemcpyHostToDevice, dev[h].st));
                { PbProfScope _prof("k_pack_results", dev[h].st, false);
                  k_pack_results<<<(hb[h].nb + 63) / 64, 64, 0, dev[h].st>>>(dev[h].children, dev[h].stats, gidx[h].p, hb[h].nb, xres.p); }
            }
            if (nhalf == 2 && hb[1].nb) { // join: the exchange runs on the first stream
                PB_CUDA_OK(cudaEventRecord(hx.fork, hx.st));
                PB_CUDA_OK(cudaStreamWaitEvent(st, hx.fork, 0));
            }
            pb_nccl_allgather(xres.p, xall.p, (size_t)nc * sizeof(XRes), st);
            static thread_local std::vector<XRes> hall;
            hall.resize((size_t)nc * sh.world);
            PB_CUDA_OK(cudaMemcpyAsync(hall.data(), xall.p, hall.size() * sizeof(XRes), cudaMemcpyDeviceToHost, st));
            PB_CUDA_OK(cudaStreamSynchronize(st));
            for (int c = 0; c < nc; c++) {
                const XRes &x = hall[(size_t)(exec[c] < 0 ? sh.rank : exec[c]) * nc + c];
                HPair *o = outs[cand[c].i];
                o->valid = true;
                o->l = HNode{x.seg[0], x.st[0]};
                o->r = HNode{x.seg[1], x.st[1]};
                o->l.owner = o->r.owner = exec[c];
                if (exec[c] < 0 && !exact) { // replicated cluster: every rank ran its own (summation-order dependent)
                    bool refused = false;    // certificate - all of them must agree on the re-evaluation
                    for (int r = 0; r < sh.world; r++) refused |= hall[(size_t)r * nc + c].seg[0].pad == PB_ROUTE_REFUSED;
                    if (refused) o->l.seg.pad = o->r.seg.pad = PB_ROUTE_REFUSED;
                }
            }
            return;
        }
        for (int h = 0; h < nhalf; h++) { // pageable destinations: each copy returns when its stream got there
            if (!hb[h].nb) continue;
            PB_CUDA_OK(cudaMemcpyAsync(hb[h].hch, dev[h].children, 2 * hb[h].nb * sizeof(PbSeg), cudaMemcpyDeviceToHost, dev[h].st));
            PB_CUDA_OK(cudaMemcpyAsync(hb[h].hst, dev[h].stats, 2 * hb[h].nb * sizeof(PbStats), cudaMemcpyDeviceToHost, dev[h].st));
        }
        for (int h = 0; h < nhalf; h++)
            if (hb[h].nb) PB_CUDA_OK(cudaStreamSynchronize(dev[h].st));
        for (int h = 0; h < nhalf; h++)
            if (hb[h].nb) shard_merge(hb[h].hst, 2 * hb[h].nb, false, weighted);
        for (int h = 0; h < nhalf; h++)
            for (int b = 0; b < hb[h].nb; b++) {
                HPair *o = outs[hb[h].map[b]];
                o->valid = true;
                o->l = HNode{hb[h].hch[2 * b], hb[h].hst[2 * b]};
                o->r = HNode{hb[h].hch[2 * b + 1], hb[h].hst[2 * b + 1]};
            }
    }

    static double benefit_of(const HNode &c, const HPair &ch) { // local.c:256-275
        return ch.valid ? c.st.dist - (ch.l.st.dist + ch.r.st.dist) : 0;
    }

    // ---- LQ (local.c:318-404) ---------------------------------------------------------
    // The reference splits the best leaf, then evaluates split_cluster() of its two children
    // before it can pick again (local.c:378-379): 2 clusters per step, K-1 dependent steps.
    // split_cluster(c) is a pure function of c's pixel set, so the ORDER in which leaves get
    // evaluated is free.  We therefore pre-evaluate the children of the leaves that will be
    // picked soonest (the current top benefits) together with the ones needed right now - one
    // batch of up to MAXB clusters per host round trip instead of two.  Picks, slot assignment
    // and termination follow the reference exactly; only the scheduling differs.
    struct Pre {
        bool have = false;
        HPair l, r; // split_cluster(children[slot].l), split_cluster(children[slot].r)
    };

    void run_lq(std::vector<HNode> &clusters, size_t K) {
        size_t len = clusters.size();
        if (len >= K) return;
        clusters.resize(K);
        std::vector<HPair> children(K);
        std::vector<Pre> pre(K);
        for (size_t i = 0; i < len; i += MAXB) { // local.c:342-345
            HNode *nn[MAXB];
            HPair *oo[MAXB];
            int c = 0;
            for (size_t j = i; j < len && c < MAXB; j++, c++) { nn[c] = &clusters[j]; oo[c] = &children[j]; }
            eval_split(nn, oo, c);
        }
        size_t i;
        for (i = len; i < K; i++) {
            // local.c:277-307 + vector.c:26-46: first maximum of d - (dl + dr)
            size_t best = 0;
            double bb = 0;
            for (size_t j = 0; j < i; j++) {
                const double b = benefit_of(clusters[j], children[j]);
                if (j == 0 || b > bb) { bb = b; best = j; }
            }
            if (bb < PB_DELTA) break; // local.c:365-370
            if (!pre[best].have) {
                // evaluate the grandchildren of `best` now, and speculatively those of the next
                // most beneficial leaves (they are the next picks unless a new leaf overtakes them)
                std::vector<size_t> cand;
                for (size_t j = 0; j < i; j++)
                    if (j != best && children[j].valid && !pre[j].have && benefit_of(clusters[j], children[j]) >= PB_DELTA)
                        cand.push_back(j);
                std::sort(cand.begin(), cand.end(), [&](size_t a, size_t b) {
                    return benefit_of(clusters[a], children[a]) > benefit_of(clusters[b], children[b]);
                });
                const size_t remaining = K - i; // picks still to make, this one included
                size_t take = std::min<size_t>({cand.size(), (size_t)MAXB / 2 - 1, remaining - 1});
                HNode *nn[MAXB];
                HPair *oo[MAXB];
                int c = 0;
                auto add = [&](size_t slot) {
                    nn[c] = &children[slot].l; oo[c++] = &pre[slot].l;
                    nn[c] = &children[slot].r; oo[c++] = &pre[slot].r;
                    pre[slot].have = true;
                };
                add(best);
                for (size_t t = 0; t < take; t++) add(cand[t]);
                eval_split(nn, oo, c);
            }
            const Pre p = pre[best];
            clusters[i] = children[best].l;    // local.c:375
            clusters[best] = children[best].r; // local.c:376
            children[i] = p.l;                 // local.c:378
            children[best] = p.r;              // local.c:379
            pre[i] = Pre{};
            pre[best] = Pre{};
        }
        clusters.resize(i);
        final_children.assign(children.begin(), children.begin() + i);
    }
    // children of the final leaves: a leaf whose split was (pre-)evaluated had its own planes
    // overwritten by its grandchildren's partition, so its members are read from its two children
    std::vector<HPair> final_children;
};

void set_timing(int slot, double ms) { g_timings[slot] = ms; }

// FP64 throughput probe (the denominator of bench.py's FP64-ALU fractions): 8 independent DFMA chains per thread
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = (double)(threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __fma_rn(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 12345.678) out[0] = s; // never true: keeps the chains alive
}

const char *k_messages[8] = {
    "Quantization successful.",
    "Internal quantization error.",
    "Image dimensions should be greater than 0.",
    "Palette size should be greater than 0.",
    "Image dimensions are too big.",
    "CUDA error (no usable sm_100 device, or out of device memory).",
    "Palette size above 50000 with KMeans refinement (kmeans_niter > 0) is not supported by patolette_b200.",
    "Saliency weights (tile_size > 0) need an image with both sides above 3 and border strips that fit: 1 <= floor(0.1 * sqrt(width * height)) < min(width, height).",
};

// Palette (K x 3 row-major, host) through one of the colour kernels.
void palette_transform(Quantizer &qz, int which, std::vector<double> &pal_rm) {
    const size_t K = pal_rm.size() / 3;
    if (!K) return;
    std::vector<double> planar(3 * K);
    for (size_t j = 0; j < K; j++)
        for (int c = 0; c < 3; c++) planar[c * K + j] = pal_rm[3 * j + c];
    DevArr<double> d;
    d.alloc(3 * K);
    qz.h2d(d.p, planar.data(), 3 * K);
    const double *src[3] = {d.p, d.p + K, d.p + 2 * K};
    double *dst[3] = {d.p, d.p + K, d.p + 2 * K};
    pb_launch_color(which, src, dst, K, qz.sm_count, qz.st);
    qz.d2h(planar.data(), d.p, 3 * K);
    qz.sync();
    for (size_t j = 0; j < K; j++)
        for (int c = 0; c < 3; c++) pal_rm[3 * j + c] = planar[c * K + j];
}

void colors_transform(Quantizer &qz, int which, size_t off = 0, size_t len = (size_t)-1) {
    if (len == (size_t)-1) len = qz.N - off;
    if (!len) return;
    const double *src[3] = {qz.col[0].p + off, qz.col[1].p + off, qz.col[2].p + off};
    double *dst[3] = {qz.col[0].p + off, qz.col[1].p + off, qz.col[2].p + off};
    pb_prof_next_bytes(48.0 * len);
    pb_launch_color(which, src, dst, len, qz.sm_count, qz.st);
}

// How the caller's buffers look.  The reference ABI is {host, planar f64, size_t map}.
struct IoSpec {
    bool device_io = false; // data / weights / map are device pointers
    int in_fmt = 0;         // 0: three f64 planes (column-major N x 3), 1: N x 3 row-major f64, 2: N x 3 row-major uint8 (/ 255 on the device)
    int map_bytes = 8;      // palette_map element: 8 (size_t), 1 or 2
    bool sharded = false;   // data / weights / map hold this rank's pixel slice only (NCCL communicator required)
    double tile_size = 0.0; // > 0 and no weights given: saliency weights from the sRGB input (patolette.pyx:411-415, row N3)
};
enum { IN_PLANAR = 0, IN_INTERLEAVED = 1, IN_U8 = 2 };

void run_patolette(size_t width, size_t height, const void *data_v, const double *weights, size_t K,
                   const patolette__QuantizationOptions *opt, double *palette, void *palette_map,
                   int *exit_code, const IoSpec &io) {
    const size_t n = width * height;
    const bool device_io = io.device_io;
    // the f32 KMeans slice sorts samples by a 16-bit assignment through per-warp class counters in shared
    // memory: beyond PB_KMEANS_MAX_K the call fails with its own code instead of silently skipping refinement
    if (opt->kmeans_niter > 0 && K > PB_KMEANS_MAX_K && n >= K) { *exit_code = -6; return; }
    if (io.sharded && !pb_nccl_active()) { *exit_code = -1; return; }
    // saliency needs the whole image on one device (raster scans, border strips): not offered for pixel slices
    const bool saliency = io.tile_size > 0.0 && weights == nullptr;
    if (saliency && io.sharded) { *exit_code = -1; return; }
    memset(g_timings, 0, sizeof g_timings);
    g_saliency_ms = 0.0;
    const long launches0 = pb_prof_launch_count();
    Quantizer qz;
    qz.sh = make_shard_ctx(n, io.sharded);
    const ShardCtx &sh = qz.sh;
    // the caller's buffers cover pixels [in_off, in_off + in_n) of the image
    const size_t in_n = sh.on ? sh.count : n, in_off = sh.on ? sh.first : 0;
    qz.init(n, weights != nullptr || saliency, sh.on ? sh.S * (size_t)sh.world : n);
    Timer total(qz.st), stage(qz.st);
    total.start();
    struct SideStream { // copies that overlap kernels of the compute stream (pinned host buffers only)
        cudaStream_t s = nullptr;
        cudaStream_t get() {
            if (!s) PB_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            return s;
        }
        ~SideStream() { if (s) cudaStreamDestroy(s); }
    } copy_stream;
    bool piped_color = false;
    const int to_space = opt->color_space == patolette__CIELuv ? PB_T_SRGB_TO_CIELUV
                         : opt->color_space == patolette__ICtCp ? PB_T_SRGB_TO_ICTCP : -1;

    stage.start(); // patolette.c:187-199: the library works on its own copy
    const cudaMemcpyKind in_kind = device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    double *const cdst[3] = {qz.col[0].p + in_off, qz.col[1].p + in_off, qz.col[2].p + in_off};
    // chunked delivery from pinned host memory: chunk c is converted / colour-transformed on the compute stream
    // while chunk c + 1 crosses PCIe.  copy(off, len) enqueues the chunk's copies on the copy stream, then
    // land(off, len) its kernels on the compute stream.
    auto piped = [&](auto copy, auto land) {
        constexpr int NCH = 8;
        const size_t per = ((in_n + NCH - 1) / NCH + 1023) & ~(size_t)1023;
        for (size_t off = 0; off < in_n; off += per) {
            const size_t len = std::min(per, in_n - off);
            copy(off, len);
            cudaEvent_t ev;
            PB_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            PB_CUDA_OK(cudaEventRecord(ev, copy_stream.get()));
            PB_CUDA_OK(cudaStreamWaitEvent(qz.st, ev, 0));
            PB_CUDA_OK(cudaEventDestroy(ev)); // released by the runtime once it has completed
            land(off, len);
        }
    };
    const bool pinned_in = !device_io && in_n >= ((size_t)1 << 20) && pb_host_is_pinned(data_v);
    const int to_space_piped = saliency ? -1 : to_space; // the saliency stage reads the sRGB planes first
    DevArr<uint8_t> rgb8;
    if (io.in_fmt == IN_U8) { // N1: uint8 RGB in, / 255 on the device
        const uint8_t *src = (const uint8_t *)data_v;
        rgb8.alloc(3 * in_n + 16);
        if (pinned_in) {
            piped_color = true;
            piped([&](size_t off, size_t len) {
                      PB_CUDA_OK(cudaMemcpyAsync(rgb8.p + 3 * off, src + 3 * off, 3 * len, in_kind, copy_stream.get()));
                  },
                  [&](size_t off, size_t len) {
                      double *const d[3] = {cdst[0] + off, cdst[1] + off, cdst[2] + off};
                      pb_prof_next_bytes(27.0 * len);
                      pb_launch_u8_to_planes(rgb8.p + 3 * off, len, d, qz.sm_count, qz.st);
                      if (to_space_piped >= 0) colors_transform(qz, to_space_piped, in_off + off, len);
                  });
            piped_color = !saliency;
        } else {
            if (device_io) PB_CUDA_OK(cudaMemcpyAsync(rgb8.p, src, 3 * in_n, in_kind, qz.st));
            else pb_copy_h2d(rgb8.p, src, 3 * in_n, qz.st);
            pb_prof_next_bytes(27.0 * in_n);
            pb_launch_u8_to_planes(rgb8.p, in_n, cdst, qz.sm_count, qz.st);
        }
    } else if (io.in_fmt == IN_INTERLEAVED) { // N x 3 row-major input: one copy, de-interleaved on the device
        const double *data = (const double *)data_v;
        DevArr<double> rgb;
        rgb.alloc(3 * in_n);
        if (device_io) PB_CUDA_OK(cudaMemcpyAsync(rgb.p, data, 3 * in_n * sizeof(double), in_kind, qz.st));
        else pb_copy_h2d(rgb.p, data, 3 * in_n * sizeof(double), qz.st);
        pb_launch_deinterleave(rgb.p, in_n, cdst, qz.sm_count, qz.st);
        qz.sync(); // rgb is released at the end of this scope
    } else if (pinned_in) {
        // pinned host planes: the copy is pipelined with the colour transform (patolette.c:201-207)
        const double *data = (const double *)data_v;
        piped_color = true;
        piped([&](size_t off, size_t len) {
                  for (int j = 0; j < 3; j++)
                      PB_CUDA_OK(cudaMemcpyAsync(cdst[j] + off, data + (size_t)j * in_n + off, len * sizeof(double), in_kind, copy_stream.get()));
              },
              [&](size_t off, size_t len) {
                  if (to_space_piped >= 0) colors_transform(qz, to_space_piped, in_off + off, len);
              });
        piped_color = !saliency;
    } else {
        const double *data = (const double *)data_v;
        for (int j = 0; j < 3 && in_n; j++) {
            if (device_io) PB_CUDA_OK(cudaMemcpyAsync(cdst[j], data + (size_t)j * in_n, in_n * sizeof(double), in_kind, qz.st));
            else pb_copy_h2d(cdst[j], data + (size_t)j * in_n, in_n * sizeof(double), qz.st);
        }
    }
    if (weights && in_n) {
        if (device_io) PB_CUDA_OK(cudaMemcpyAsync(qz.wgt.p + in_off, weights, in_n * sizeof(double), in_kind, qz.st));
        else pb_copy_h2d(qz.wgt.p + in_off, weights, in_n * sizeof(double), qz.st);
    }
    set_timing(1, stage.stop());

    stage.start(); // patolette.c:201-207
    if (saliency) { // patolette.pyx:411-415: the weights come from the sRGB image, before the colour transform
        Timer sal(qz.st);
        sal.start();
        const double *planes[3] = {qz.col[0].p, qz.col[1].p, qz.col[2].p};
        const int rc = pb_saliency_weights(planes, width, height, io.tile_size, qz.wgt.p, qz.sm_count, qz.st);
        g_saliency_ms = sal.stop();
        if (rc != 0) { *exit_code = rc; return; }
    }
    if (!piped_color && to_space >= 0) colors_transform(qz, to_space, in_off, in_n);
    if (sh.on) { // every rank transformed its slice: all-gather the planes over NVLink (in place)
        const size_t sbytes = sh.S * sizeof(double);
        for (int j = 0; j < 3; j++) pb_nccl_allgather(qz.col[j].p + (size_t)sh.rank * sh.S, qz.col[j].p, sbytes, qz.st);
        if (weights) pb_nccl_allgather(qz.wgt.p + (size_t)sh.rank * sh.S, qz.wgt.p, sbytes, qz.st);
    }
    set_timing(2, stage.stop());
    if (opt->verbose) printf("patolette ======== Palette generation \n");

    qz.init_tree();
    stage.start();
    std::vector<HNode> clusters;
    size_t count = qz.run_gq(K, clusters); // patolette.c:213
    set_timing(3, stage.stop());
    if (count == 0) { *exit_code = -1; return; }
    if (opt->verbose) printf("patolette ======== Base cluster count: %zu\n", count);
    stage.start();
    qz.run_lq(clusters, K); // patolette.c:231
    set_timing(4, stage.stop());
    count = clusters.size();

    std::vector<double> pal(3 * count); // palette/create.c:11-33
    for (size_t j = 0; j < count; j++)
        for (int c = 0; c < 3; c++) pal[3 * j + c] = clusters[j].st.mean[c];

    if (opt->kmeans_niter > 0) { // patolette.c:248-261 -> refine.c:165-221
        if (opt->verbose) printf("patolette ======== KMeans refinement\n");
        stage.start();
        const double *planes[3] = {qz.col[0].p, qz.col[1].p, qz.col[2].p};
        size_t ms = opt->kmeans_max_samples > 65536 ? opt->kmeans_max_samples : 65536; // refine.c:21,87
        pb_kmeans_refine(planes, qz.weighted ? qz.wgt.p : nullptr, n, pal, opt->kmeans_niter,
                         (int)(ms / count), qz.sm_count, qz.st, &qz.launches);
        set_timing(5, stage.stop());
    }

    if (!opt->palette_only) {
        // the caller's map covers the same pixels as its input; the device map is always whole-image size_t
        const size_t out_n = in_n, out_off = in_off;
        const size_t mb = (size_t)io.map_bytes;
        DevArr<unsigned long long> dmap;
        dmap.alloc(n);
        DevArr<uint8_t> narrow; // uint8 / uint16 indices of the caller's pixels
        if (mb != 8) narrow.alloc(out_n * mb + 16);
        bool map_sent = false;
        if (opt->dither) { // patolette.c:268-299
            if (opt->verbose) printf("patolette ======== Dithering\n");
            stage.start();
            const int t = opt->color_space == patolette__CIELuv  ? PB_T_CIELUV_TO_REC2020
                          : opt->color_space == patolette__ICtCp ? PB_T_ICTCP_TO_REC2020
                                                                 : PB_T_SRGB_TO_REC2020;
            colors_transform(qz, t);
            palette_transform(qz, t, pal);
            const double *planes[3] = {qz.col[0].p, qz.col[1].p, qz.col[2].p};
            // riemersma.c:452-456: a 1x1 image is never dithered; the map keeps the caller's bytes
            if (palette_map && n == 1 && out_n == 1) {
                PB_CUDA_OK(cudaMemsetAsync(dmap.p, 0, sizeof(size_t), qz.st));
                PB_CUDA_OK(cudaMemcpyAsync(dmap.p, palette_map, mb, in_kind, qz.st)); // (little endian)
            }
            // (an image-sharded run splits the speculative chains of the walk over the ranks; the recurrence's
            //  boundary repair stays replicated)
            const PbDitherShard ds{sh.rank, sh.world, out_off, out_n};
            pb_dither_riemersma(planes, width, height, pal, dmap.p, qz.sm_count, qz.st, &qz.launches, sh.on ? &ds : nullptr);
            palette_transform(qz, PB_T_REC2020_TO_SRGB, pal);
            set_timing(7, stage.stop());
        } else { // patolette.c:300-324
            if (opt->verbose) printf("patolette ======== NN mapping\n");
            stage.start();
            if (opt->color_space == patolette__CIELuv) {
                colors_transform(qz, PB_T_CIELUV_TO_ICTCP, out_off, out_n); // (a sharded run maps its own pixels only)
                palette_transform(qz, PB_T_CIELUV_TO_ICTCP, pal);
            }
            DevArr<double> dpal;
            dpal.alloc(3 * count);
            qz.h2d(dpal.p, pal.data(), 3 * count);
            const double *planes[3] = {qz.col[0].p + out_off, qz.col[1].p + out_off, qz.col[2].p + out_off};
            // exact 1-NN through per-cell candidate lists when the brute force would be FP64-bound (pb_nngrid.cu)
            DevArr<char> nngrid;
            const bool use_grid = g_nn_grid && count >= 32 && count <= 4096 && out_n >= ((size_t)1 << 18);
            if (use_grid) {
                nngrid.alloc(pb_nngrid_scratch_bytes((int)count));
                pb_launch_nngrid_build(planes, out_n, dpal.p, (int)count, nngrid.p, qz.sm_count, qz.st);
            }
            auto assign = [&](const double *const pl[3], size_t len, unsigned long long *out) {
                if (!len) return;
                pb_prof_next_bytes(32.0 * len);
                if (use_grid) pb_launch_nearest_grid(pl, len, dpal.p, (int)count, nngrid.p, out, qz.sm_count, qz.st);
                else pb_launch_nearest(pl, len, dpal.p, (int)count, out, qz.sm_count, qz.st);
            };
            if (!device_io && out_n >= ((size_t)1 << 20) && pb_host_is_pinned(palette_map)) {
                // pinned destination: the map goes home chunk by chunk while the next chunk is assigned
                constexpr int NCH = 4;
                const size_t per = ((out_n + NCH - 1) / NCH + 1023) & ~(size_t)1023;
                for (size_t off = 0; off < out_n; off += per) {
                    const size_t len = std::min(per, out_n - off);
                    const double *pl[3] = {planes[0] + off, planes[1] + off, planes[2] + off};
                    assign(pl, len, dmap.p + out_off + off);
                    const void *from = dmap.p + out_off + off;
                    if (mb != 8) {
                        pb_launch_narrow_map(dmap.p + out_off + off, len, narrow.p + off * mb, (int)mb, qz.sm_count, qz.st);
                        from = narrow.p + off * mb;
                    }
                    cudaEvent_t ev;
                    PB_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                    PB_CUDA_OK(cudaEventRecord(ev, qz.st));
                    PB_CUDA_OK(cudaStreamWaitEvent(copy_stream.get(), ev, 0));
                    PB_CUDA_OK(cudaEventDestroy(ev));
                    PB_CUDA_OK(cudaMemcpyAsync((char *)palette_map + off * mb, from, len * mb, cudaMemcpyDeviceToHost, copy_stream.get()));
                }
                map_sent = true;
            } else {
                assign(planes, out_n, dmap.p + out_off);
            }
            qz.sync();
            // patolette.c:322-323, applied whatever the colour space was (reference bug B1)
            palette_transform(qz, PB_T_ICTCP_TO_REC2020, pal);
            palette_transform(qz, PB_T_REC2020_TO_SRGB, pal);
            set_timing(6, stage.stop());
        }
        stage.start();
        if (map_sent) PB_CUDA_OK(cudaStreamSynchronize(copy_stream.get())); // the tail of the chunked copy
        else if (out_n) {
            const void *from = dmap.p + out_off;
            if (mb != 8) {
                pb_launch_narrow_map(dmap.p + out_off, out_n, narrow.p, (int)mb, qz.sm_count, qz.st);
                from = narrow.p;
            }
            if (device_io) PB_CUDA_OK(cudaMemcpyAsync(palette_map, from, out_n * mb, cudaMemcpyDeviceToDevice, qz.st));
            else pb_copy_d2h(palette_map, from, out_n * mb, qz.st);
        }
        qz.sync();
        set_timing(8, stage.stop());
    }
    for (size_t j = 0; j < K * 3; j++) palette[j] = -1.0; // patolette.c:328-330
    for (int c = 0; c < 3; c++)
        for (size_t j = 0; j < count; j++) palette[K * c + j] = pal[3 * j + c];
    set_timing(0, total.stop());
    g_timings[9] = (double)(pb_prof_launch_count() - launches0);
    *exit_code = 0;
}

// Every entry point that runs the pipeline: one call at a time, and nothing but an exit code crosses the C ABI.
// Before buffers go back to the cache on an error path the device is drained (kernels of the failed call may
// still be running on the side streams).
template <typename F>
void guarded_run(int *exit_code, F &&body) {
    std::lock_guard<std::mutex> lk(g_call_mu);
    try {
        body();
    } catch (const pb_cuda_error &) {
        cudaGetLastError();
        cudaDeviceSynchronize();
        cudaGetLastError();
        *exit_code = -5;
    } catch (const pb_shard_error &) {
        cudaDeviceSynchronize();
        *exit_code = -1;
    } catch (...) { // std::bad_alloc, std::system_error from a host thread, std::length_error, ...
        cudaDeviceSynchronize();
        cudaGetLastError();
        *exit_code = -1;
    }
}
template <typename F>
int guarded_stage(F &&body) {
    std::lock_guard<std::mutex> lk(g_call_mu);
    try {
        return body();
    } catch (const pb_cuda_error &e) {
        cudaGetLastError();
        cudaDeviceSynchronize();
        cudaGetLastError();
        return -(int)e.code;
    } catch (...) {
        cudaDeviceSynchronize();
        cudaGetLastError();
        return -1;
    }
}

} // namespace

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

void patolette(size_t width, size_t height, const double *data, const double *weights, size_t palette_size,
               const patolette__QuantizationOptions *options, double *palette, size_t *palette_map,
               int *exit_code) {
    *exit_code = 0; // validate_arguments, patolette.c:61-95 (same order)
    if (width * height == 0) { *exit_code = -2; return; }
    if (palette_size < 1) { *exit_code = -3; return; }
    if (width * height > (size_t)40000 * 40000) { *exit_code = -4; return; }
    guarded_run(exit_code, [&] {
        run_patolette(width, height, data, weights, palette_size, options, palette, palette_map, exit_code, IoSpec{});
    });
}

void patolette_b200_device(size_t width, size_t height, const double *d_data, const double *d_weights,
                           size_t palette_size, const patolette__QuantizationOptions *options, double *palette,
                           size_t *d_palette_map, int *exit_code) {
    *exit_code = 0;
    if (width * height == 0) { *exit_code = -2; return; }
    if (palette_size < 1) { *exit_code = -3; return; }
    if (width * height > (size_t)40000 * 40000) { *exit_code = -4; return; }
    guarded_run(exit_code, [&] {
        IoSpec io;
        io.device_io = true;
        run_patolette(width, height, d_data, d_weights, palette_size, options, palette, d_palette_map, exit_code, io);
    });
}

void patolette_b200_interleaved(size_t width, size_t height, const double *rgb, const double *weights,
                                size_t palette_size, const patolette__QuantizationOptions *options, double *palette,
                                size_t *palette_map, int *exit_code) {
    *exit_code = 0;
    if (width * height == 0) { *exit_code = -2; return; }
    if (palette_size < 1) { *exit_code = -3; return; }
    if (width * height > (size_t)40000 * 40000) { *exit_code = -4; return; }
    guarded_run(exit_code, [&] {
        IoSpec io;
        io.in_fmt = IN_INTERLEAVED;
        run_patolette(width, height, rgb, weights, palette_size, options, palette, palette_map, exit_code, io);
    });
}

void patolette_b200_u8(size_t width, size_t height, const uint8_t *rgb, const double *weights, size_t palette_size,
                       const patolette__QuantizationOptions *options, double *palette, void *palette_map,
                       int map_bytes, int device_io, int *exit_code) {
    *exit_code = 0;
    if (width * height == 0) { *exit_code = -2; return; }
    if (palette_size < 1) { *exit_code = -3; return; }
    if (width * height > (size_t)40000 * 40000) { *exit_code = -4; return; }
    // the index type must hold palette_size - 1 (no silent truncation)
    if (!(map_bytes == 1 || map_bytes == 2 || map_bytes == 8) || (map_bytes == 1 && palette_size > 256) ||
        (map_bytes == 2 && palette_size > 65536)) { *exit_code = -3; return; }
    guarded_run(exit_code, [&] {
        IoSpec io;
        io.in_fmt = IN_U8;
        io.map_bytes = map_bytes;
        io.device_io = device_io != 0;
        run_patolette(width, height, rgb, weights, palette_size, options, palette, palette_map, exit_code, io);
    });
}

void patolette_b200_quantize(size_t width, size_t height, const void *colors, int in_fmt, double tile_size,
                             size_t palette_size, const patolette__QuantizationOptions *options, double *palette,
                             void *palette_map, int map_bytes, int device_io, int *exit_code) {
    *exit_code = 0;
    if (width * height == 0) { *exit_code = -2; return; }
    if (palette_size < 1) { *exit_code = -3; return; }
    if (width * height > (size_t)40000 * 40000) { *exit_code = -4; return; }
    if (in_fmt < IN_PLANAR || in_fmt > IN_U8 || !(tile_size >= 0.0)) { *exit_code = -1; return; }
    if (!(map_bytes == 1 || map_bytes == 2 || map_bytes == 8) || (map_bytes == 1 && palette_size > 256) ||
        (map_bytes == 2 && palette_size > 65536)) { *exit_code = -3; return; }
    guarded_run(exit_code, [&] {
        IoSpec io;
        io.in_fmt = in_fmt;
        io.map_bytes = map_bytes;
        io.device_io = device_io != 0;
        io.tile_size = tile_size;
        run_patolette(width, height, colors, nullptr, palette_size, options, palette, palette_map, exit_code, io);
    });
}

// get_weights / mbd of the reference's wrapper as stages (patolette.pyx:203-313, :153-201)
static int saliency_stage(size_t width, size_t height, const double *planar, double tile_size, void *out, int device_io, bool mbd_only) {
    const size_t n = width * height;
    if (n == 0 || !planar || !out || (!mbd_only && !(tile_size > 0.0))) return -1;
    return guarded_stage([&]() -> int {
        Quantizer qz;
        qz.init(n, !mbd_only);
        const cudaMemcpyKind in_kind = device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        for (int j = 0; j < 3; j++) {
            if (device_io) PB_CUDA_OK(cudaMemcpyAsync(qz.col[j].p, planar + (size_t)j * n, n * sizeof(double), in_kind, qz.st));
            else pb_copy_h2d(qz.col[j].p, planar + (size_t)j * n, n * sizeof(double), qz.st);
        }
        const double *planes[3] = {qz.col[0].p, qz.col[1].p, qz.col[2].p};
        Timer sal(qz.st);
        sal.start();
        int rc;
        DevArr<float> dmap;
        if (mbd_only) {
            dmap.alloc(n);
            rc = pb_saliency_mbd(planes, width, height, dmap.p, qz.sm_count, qz.st);
        } else {
            rc = pb_saliency_weights(planes, width, height, tile_size, qz.wgt.p, qz.sm_count, qz.st);
        }
        g_saliency_ms = sal.stop();
        if (rc != 0) return rc;
        const void *from = mbd_only ? (const void *)dmap.p : (const void *)qz.wgt.p;
        const size_t bytes = n * (mbd_only ? sizeof(float) : sizeof(double));
        if (device_io) PB_CUDA_OK(cudaMemcpyAsync(out, from, bytes, cudaMemcpyDeviceToDevice, qz.st));
        else pb_copy_d2h(out, from, bytes, qz.st);
        qz.sync();
        return 0;
    });
}
int patolette_b200_saliency_weights(size_t width, size_t height, const double *planar, double tile_size, double *weights,
                                    int device_io) {
    return saliency_stage(width, height, planar, tile_size, weights, device_io, false);
}
int patolette_b200_saliency_mbd(size_t width, size_t height, const double *planar, float *distance, int device_io) {
    return saliency_stage(width, height, planar, 0.0, distance, device_io, true);
}
double patolette_b200_last_saliency_ms(void) { return g_saliency_ms; }

// N2 stage: n symmetric 3 x 3 solves, on the host (on_device = 0, no GPU involved) or by k_eigen3
int patolette_b200_eigen3(const double *a9, size_t n, double *w3, double *z9, int *info, int on_device) {
    if (!a9 || !w3 || !z9 || n > (size_t)1 << 28) return -1;
    if (!on_device) {
        for (size_t i = 0; i < n; i++) {
            double a[9];
            memcpy(a, a9 + 9 * i, sizeof a);
            const int rc = pb_eig::dsyev3(a, w3 + 3 * i);
            memcpy(z9 + 9 * i, a, sizeof a);
            if (info) info[i] = rc;
        }
        return 0;
    }
    return guarded_stage([&]() -> int {
        Quantizer qz;
        qz.init(0, false);
        DevArr<double> da, dw, dz;
        DevArr<int> di;
        da.alloc(9 * n + 1); dw.alloc(3 * n + 1); dz.alloc(9 * n + 1); di.alloc(n + 1);
        qz.h2d(da.p, a9, 9 * n);
        pb_launch_eigen3(da.p, (int)n, dw.p, dz.p, di.p, qz.st);
        qz.d2h(w3, dw.p, 3 * n);
        qz.d2h(z9, dz.p, 9 * n);
        if (info) qz.d2h(info, di.p, n);
        qz.sync();
        return 0;
    });
}

void patolette_b200_sharded(size_t width, size_t height, const double *slice, const double *weights_slice,
                            size_t palette_size, const patolette__QuantizationOptions *options, double *palette,
                            size_t *map_slice, int device_io, int *exit_code) {
    *exit_code = 0;
    if (width * height == 0) { *exit_code = -2; return; }
    if (palette_size < 1) { *exit_code = -3; return; }
    if (width * height > (size_t)40000 * 40000) { *exit_code = -4; return; }
    guarded_run(exit_code, [&] {
        IoSpec io;
        io.sharded = true;
        io.device_io = device_io != 0;
        run_patolette(width, height, slice, weights_slice, palette_size, options, palette, map_slice, exit_code, io);
    });
}

int patolette_b200_shard_range(size_t n_pixels, int rank, int world, size_t *first, size_t *count) {
    if (world < 1 || rank < 0 || rank >= world) return -1;
    const size_t S = shard_stride(n_pixels, world);
    const size_t f = std::min(n_pixels, (size_t)rank * S);
    if (first) *first = f;
    if (count) *count = std::min(S, n_pixels - f);
    return 0;
}

int patolette_b200_comm_unique_id(char *id128) { return id128 ? pb_nccl_unique_id(id128) : -1; }

int patolette_b200_comm_init(int rank, int world, const char *id128) {
    std::lock_guard<std::mutex> lk(g_call_mu);
    if (cudaSetDevice(g_device) != cudaSuccess) return -5;
    return pb_nccl_init(rank, world, id128);
}

void patolette_b200_comm_destroy(void) {
    std::lock_guard<std::mutex> lk(g_call_mu);
    pb_nccl_destroy();
}

int patolette_b200_comm_info(int *rank, int *world, int *nccl_version) {
    if (rank) *rank = pb_nccl_rank();
    if (world) *world = pb_nccl_world();
    if (nccl_version) *nccl_version = pb_nccl_version();
    return pb_nccl_active() ? 1 : 0;
}

int patolette_b200_ordered_counts(unsigned long long *out2, int reset) {
    return guarded_stage([&]() -> int {
        pb_ordered_counts(out2, reset != 0);
        return 0;
    });
}

int patolette_b200_split_counts(unsigned long long *out4, int reset) {
    return guarded_stage([&]() -> int {
        pb_certify_counts(out4, reset != 0);
        out4[2] = g_split_redone;
        if (reset) g_split_redone = 0;
        return 0;
    });
}

int patolette_b200_gq_cuts(const double *bucket_sums, const unsigned int *class_start, size_t palette_size, size_t *cuts16) {
    // host only (no CUDA): the GQ dynamic programme on a table of per-bucket sums, for the CPU tests
    static thread_local CellMoments m;
    build_cell_moments(bucket_sums, class_start, m);
    size_t q[16] = {0};
    const size_t cells = principal_quantizer(palette_size, m, q);
    for (int i = 0; i < 16; i++) cuts16[i] = q[i];
    return (int)cells;
}

int patolette_b200_set_sharding(int rank, int world, patolette_b200_allgather_fn allgather, void *user) {
    if (world < 1 || rank < 0 || rank >= world || (world > 1 && !allgather)) return -1;
    g_shard_rank = rank;
    g_shard_world = world;
    g_shard_allgather = world > 1 ? allgather : nullptr;
    g_shard_user = user;
    return 0;
}

int patolette_b200_set_option(const char *name, long long value) {
    if (!name) return -1;
    if (!strcmp(name, "dump_cap")) { pb_ordered_set_dump_cap(value); return 0; }
    if (!strcmp(name, "scatter_cta")) { pb_scatter_set_cta(value != 0); return 0; }
    if (!strcmp(name, "sorted_payload")) { g_sorted_payload = value != 0; return 0; }
    if (!strcmp(name, "gq_chain_cta")) { pb_chain_set_gq_cta(value != 0); return 0; }
    if (!strcmp(name, "raw_moments")) { pb_ordered_set_raw_moments(value != 0); return 0; }
    if (!strcmp(name, "prefix_slabs")) { pb_ordered_set_prefix_slabs((int)value); return 0; }
    if (!strcmp(name, "fused_pass")) { pb_ordered_set_fused(value != 0); return 0; }
    if (!strcmp(name, "fast_summary")) { pb_ordered_set_fast(value != 0); return 0; }
    if (!strcmp(name, "prof_timeline")) { pb_prof_set_timeline(value != 0); return 0; }
    if (!strcmp(name, "split_certify")) { if (value < 0 || value > 2) return -1; g_split_certify = (int)value; return 0; }
    if (!strcmp(name, "overlap")) { g_overlap_override = (int)value; return 0; }
    if (!strcmp(name, "nn_grid")) { g_nn_grid = value != 0; return 0; }
    if (!strcmp(name, "gq_threads")) { g_gq_threads = (int)value; return 0; }
    if (!strcmp(name, "gq_full_table")) { g_gq_full_table = value != 0; return 0; }
    if (!strcmp(name, "dither_grid")) { pb_dither_set_grid(value != 0); return 0; }
    if (!strcmp(name, "dither_subwarp")) { pb_dither_set_subwarp(value != 0); return 0; }
    if (!strcmp(name, "dither_tiles")) { pb_dither_set_tiles(value != 0); return 0; }
    if (!strcmp(name, "dither_one_wave")) { pb_dither_set_one_wave(value != 0); return 0; }
    if (!strcmp(name, "allow_jacobi")) { pb_lapack_allow_jacobi(value != 0); return 0; }
    if (!strcmp(name, "host_lapack")) { pb_lapack_use_host(value != 0); return 0; }
    return -1;
}

int patolette_b200_ordered_chain_debug(unsigned long long *out35, int reset) {
    return guarded_stage([&]() -> int {
        pb_ordered_chain_debug(out35, reset != 0);
        return 0;
    });
}

size_t patolette_b200_release_cache(void) {
    std::lock_guard<std::mutex> lk(g_call_mu);
    pb_dither_release_cache();
    const size_t held = pb_pool_cached_bytes();
    pb_pool_release_all();
    return held;
}

int patolette_b200_set_stream(void *cuda_stream, int enable) {
    g_user_stream = (cudaStream_t)cuda_stream;
    g_use_user_stream = enable != 0;
    return 0;
}

int patolette_b200_profile_enable(int on) {
    pb_prof_enable(on != 0);
    if (on) pb_prof_reset();
    return 0;
}

size_t patolette_b200_profile_json(char *buf, size_t cap) {
    std::string js = pb_prof_json();
    if (buf && cap) {
        size_t m = js.size() < cap - 1 ? js.size() : cap - 1;
        memcpy(buf, js.data(), m);
        buf[m] = 0;
    }
    return js.size() + 1;
}

size_t patolette_b200_profile_timeline(char *buf, size_t cap) {
    std::string js = pb_prof_timeline_text();
    if (buf && cap) {
        size_t m = js.size() < cap - 1 ? js.size() : cap - 1;
        memcpy(buf, js.data(), m);
        buf[m] = 0;
    }
    return js.size() + 1;
}

const char *get_patolette_exit_code_info_message(int exit_code) {
    int i = -exit_code;
    if (i < 0 || i > 7) i = 1;
    return k_messages[i];
}

patolette__QuantizationOptions *patolette_create_default_options(void) { // patolette.c:107-119
    patolette__QuantizationOptions *o = (patolette__QuantizationOptions *)malloc(sizeof *o);
    o->dither = true;
    o->palette_only = false;
    o->color_space = patolette__ICtCp;
    o->kmeans_niter = 32;
    o->kmeans_max_samples = 512 * 512;
    o->verbose = false;
    return o;
}

int patolette_b200_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return -(int)e;
    g_device = device;
    return 0;
}

int patolette_b200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    return e == cudaSuccess ? n : -(int)e;
}

void patolette_b200_set_lapack(const char *path) { pb_lapack_set_path(path); }
const char *patolette_b200_lapack_source(void) { return pb_lapack_source(); }

double patolette_b200_fp64_peak(void) {
    std::lock_guard<std::mutex> lk(g_call_mu);
    try {
        PB_CUDA_OK(cudaSetDevice(g_device));
        int sms = 0;
        PB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_device));
        double *d = (double *)pb_pool_alloc(64);
        cudaEvent_t e0, e1;
        PB_CUDA_OK(cudaEventCreate(&e0));
        PB_CUDA_OK(cudaEventCreate(&e1));
        const int iters = 1 << 14, grid = sms * 8;
        double best = 0;
        for (int rep = 0; rep < 4; rep++) { // the first repetition warms up
            PB_CUDA_OK(cudaEventRecord(e0, 0));
            k_fp64_peak<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
            PB_CUDA_OK(cudaEventRecord(e1, 0));
            PB_CUDA_OK(cudaEventSynchronize(e1));
            float ms = 0;
            PB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
            const double tf = 2.0 * 8 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
            if (rep && tf > best) best = tf;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        pb_pool_free(d);
        return best;
    } catch (const pb_cuda_error &e) {
        cudaGetLastError();
        return -(double)e.code;
    }
}

int patolette_b200_last_timings(double *out10) {
    memcpy(out10, g_timings, sizeof g_timings);
    return 0;
}

int patolette_b200_color_transform(int which, double *planar, size_t n) {
    return guarded_stage([&]() -> int {
        Quantizer qz;
        qz.init(n, false);
        for (int j = 0; j < 3; j++) qz.h2d(qz.col[j].p, planar + (size_t)j * n, n);
        colors_transform(qz, which);
        for (int j = 0; j < 3; j++) qz.d2h(planar + (size_t)j * n, qz.col[j].p, n);
        qz.sync();
        return 0;
    });
}

int patolette_b200_pow(const double *x, double y, double *out, size_t n) {
    return guarded_stage([&]() -> int {
        Quantizer qz;
        qz.init(n, false);
        qz.h2d(qz.col[0].p, x, n);
        pb_launch_pow(qz.col[0].p, y, qz.col[1].p, n, qz.sm_count, qz.st);
        qz.d2h(out, qz.col[1].p, n);
        qz.sync();
        return 0;
    });
}

int patolette_b200_quantize_clusters(const double *planar, size_t n, const double *weights, size_t palette_size,
                                     uint32_t *labels, double *centers, size_t *count, size_t *gq_count) {
    return guarded_stage([&]() -> int {
        Quantizer qz;
        qz.init(n, weights != nullptr);
        for (int j = 0; j < 3; j++) qz.h2d(qz.col[j].p, planar + (size_t)j * n, n);
        if (weights) qz.h2d(qz.wgt.p, weights, n);
        qz.init_tree();
        std::vector<HNode> clusters;
        size_t c = qz.run_gq(palette_size, clusters);
        if (gq_count) *gq_count = c;
        if (c == 0) return -1;
        qz.run_lq(clusters, palette_size);
        *count = clusters.size();
        if (centers)
            for (size_t j = 0; j < clusters.size(); j++)
                for (int k = 0; k < 3; k++) centers[3 * j + k] = clusters[j].st.mean[k];
        if (labels) {
            DevArr<PbSeg> dsegs;
            DevArr<uint32_t> dlab;
            std::vector<PbSeg> hs;
            uint32_t max_n = 0;
            for (size_t j = 0; j < clusters.size(); j++) {
                // a leaf whose split was evaluated is read through its two children (see run_lq)
                const bool via_children = j < qz.final_children.size() && qz.final_children[j].valid;
                PbSeg parts[2] = {via_children ? qz.final_children[j].l.seg : clusters[j].seg,
                                  via_children ? qz.final_children[j].r.seg : PbSeg{}};
                for (int t = 0; t < (via_children ? 2 : 1); t++) {
                    parts[t].pad = (uint32_t)j;
                    max_n = std::max(max_n, parts[t].n);
                    hs.push_back(parts[t]);
                }
            }
            dsegs.alloc(hs.size());
            dlab.alloc(n);
            qz.h2d(dsegs.p, hs.data(), hs.size());
            pb_launch_labels(qz.bufs, dsegs.p, (int)hs.size(), max_n, dlab.p, qz.sm_count, qz.st);
            qz.d2h(labels, dlab.p, n);
            qz.sync();
        }
        return 0;
    });
}

int patolette_b200_nearest(const double *planar, size_t n, const double *palette_rm, size_t K, size_t *map) {
    return guarded_stage([&]() -> int {
        Quantizer qz;
        qz.init(n, false);
        for (int j = 0; j < 3; j++) qz.h2d(qz.col[j].p, planar + (size_t)j * n, n);
        DevArr<double> dpal;
        DevArr<unsigned long long> dmap;
        dpal.alloc(3 * K);
        dmap.alloc(n);
        qz.h2d(dpal.p, palette_rm, 3 * K);
        const double *planes[3] = {qz.col[0].p, qz.col[1].p, qz.col[2].p};
        DevArr<char> nngrid;
        if (g_nn_grid && K >= 32 && K <= 4096 && n >= ((size_t)1 << 18)) { // same route selection as run_patolette
            nngrid.alloc(pb_nngrid_scratch_bytes((int)K));
            pb_launch_nngrid_build(planes, n, dpal.p, (int)K, nngrid.p, qz.sm_count, qz.st);
            pb_launch_nearest_grid(planes, n, dpal.p, (int)K, nngrid.p, dmap.p, qz.sm_count, qz.st);
        } else {
            pb_launch_nearest(planes, n, dpal.p, (int)K, dmap.p, qz.sm_count, qz.st);
        }
        qz.d2h((unsigned long long *)map, dmap.p, n);
        qz.sync();
        return 0;
    });
}

int patolette_b200_kmeans(const float *x, size_t n, size_t K, float *centers, const float *w, int niter,
                          int max_points_per_centroid) {
    // Mirrors faiss kmeans_clustering (Clustering.cpp:587-603) on host-provided f32 samples by
    // widening them to the f64 planes the pipeline keeps on the device ((float)(double)f == f).
    return guarded_stage([&]() -> int {
        if (n < K) return -1;
        Quantizer qz;
        qz.init(n, w != nullptr);
        std::vector<double> tmp(n);
        for (int j = 0; j < 3; j++) {
            for (size_t i = 0; i < n; i++) tmp[i] = (double)x[3 * i + j];
            qz.h2d(qz.col[j].p, tmp.data(), n);
            qz.sync();
        }
        if (w) {
            for (size_t i = 0; i < n; i++) tmp[i] = (double)w[i];
            qz.h2d(qz.wgt.p, tmp.data(), n);
            qz.sync();
        }
        std::vector<double> pal(3 * K);
        for (size_t j = 0; j < 3 * K; j++) pal[j] = (double)centers[j];
        const double *planes[3] = {qz.col[0].p, qz.col[1].p, qz.col[2].p};
        pb_kmeans_refine(planes, w ? qz.wgt.p : nullptr, n, pal, niter, max_points_per_centroid, qz.sm_count, qz.st,
                         &qz.launches);
        for (size_t j = 0; j < 3 * K; j++) centers[j] = (float)pal[j];
        return 0;
    });
}

int patolette_b200_dither(const double *planar, size_t width, size_t height, const double *palette_rm, size_t K,
                          size_t *map) {
    return guarded_stage([&]() -> int {
        const size_t n = width * height;
        Quantizer qz;
        qz.init(n, false);
        for (int j = 0; j < 3; j++) qz.h2d(qz.col[j].p, planar + (size_t)j * n, n);
        DevArr<unsigned long long> dmap;
        dmap.alloc(n);
        bool map_sent = false;
        qz.h2d(dmap.p, (const unsigned long long *)map, n);
        std::vector<double> pal(palette_rm, palette_rm + 3 * K);
        const double *planes[3] = {qz.col[0].p, qz.col[1].p, qz.col[2].p};
        pb_dither_riemersma(planes, width, height, pal, dmap.p, qz.sm_count, qz.st, &qz.launches);
        qz.d2h((unsigned long long *)map, dmap.p, n);
        qz.sync();
        return 0;
    });
}

} // extern "C"
