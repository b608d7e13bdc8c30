// pb_kernels.h - host-callable launchers of the patolette_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "pb_common.cuh"

// Colour transforms (pb_color.cu).  Values mirror oracle ORC_T_* for tests.
enum {
    PB_T_SRGB_TO_ICTCP = 0,
    PB_T_SRGB_TO_CIELUV = 1,
    PB_T_ICTCP_TO_REC2020 = 2,
    PB_T_CIELUV_TO_REC2020 = 3,
    PB_T_SRGB_TO_REC2020 = 4,
    PB_T_REC2020_TO_SRGB = 5,
    PB_T_CIELUV_TO_ICTCP = 6, // CIELuv -> Rec2020 -> sRGB -> ICtCp (patolette.c:305-314), fused
};
void pb_launch_color(int which, const double *const src[3], double *const dst[3], size_t n,
                     int sm_count, cudaStream_t st);
void pb_launch_pow(const double *x, double y, double *out, size_t n, int sm_count, cudaStream_t st);
void pb_launch_deinterleave(const double *d_rgb, size_t n, double *const dst[3], int sm_count, cudaStream_t st);
// N x 3 interleaved uint8 -> planes of value / 255 (f64, IEEE division); size_t map -> uint8 / uint16 indices
void pb_launch_u8_to_planes(const uint8_t *d_rgb, size_t n, double *const dst[3], int sm_count, cudaStream_t st);
// batch of 3 x 3 symmetric eigen solves, LAPACK dsyev-faithful (pb_eigen.cu / pb_dsyev3.h)
void pb_launch_eigen3(const double *a9, int n, double *w3, double *z9, int *info, cudaStream_t st);
void pb_launch_narrow_map(const unsigned long long *d_map, size_t n, void *d_out, int bytes, int sm_count, cudaStream_t st);

// ---- ordered-sum kernels, pb_ordered.cu / pb_chain.cu ----------------------------------
// Every reference statistic is a left-to-right f64 sum in ascending pixel order; these
// kernels reproduce that order bit for bit (see pb_chain.cu for how).
// max_n = largest segment of the batch (grid width); block tables are packed: segment s owns
// blocks [s.bbase, s.bbase + blocks(s.n)) and total_blocks bounds the table.  With d_scratch
// (pb_ordered_scratch_bytes(total_blocks)) large segments take the speculative block-summary
// path, otherwise (and for small segments) the sums run as in-warp speculative replays only.
uint32_t pb_ordered_blocks(uint32_t n);
size_t pb_ordered_scratch_bytes(size_t total_blocks);
// {blocks accepted from summaries, blocks replayed sequentially} per chain since the last reset
void pb_ordered_counts(unsigned long long out[16], bool reset);
void pb_ordered_chain_debug(unsigned long long out[35], bool reset);
// test knob: cap on the term-dump slots of a pass (-1 = default); blocks beyond it are replayed from the planes
void pb_ordered_set_dump_cap(long long slots);
// test knob: block-uniform fast summaries (default) or the per-element summary kernel for every block
void pb_ordered_set_fast(bool on);
// test knob: single-pass fused summaries (default) or the separate blocksum + prefix + summary kernels
void pb_ordered_set_fused(bool on);
// test knob: block-sum scans of long segments shared by several CTAs (default) or one CTA per chain
void pb_ordered_set_prefix_slabs(int mode); // 0 off, 1 auto, >= 2 forced slab count
// test knob: centred-pass block sums derived from the mean pass's raw moments (default) or summed from the pixels
void pb_ordered_set_raw_moments(bool on);
// cmask: bit c set = chain c is summed by this call (chain-sharded multi-GPU runs split the chains over the
// ranks; the other fields of PbStats are left as they are).  raw_mean: leave sum(c_j * w) unscaled in mean[].
void pb_launch_pass_mean(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                         uint32_t total_blocks, bool weighted, PbStats *d_stats, void *d_scratch,
                         size_t scratch_bytes, cudaStream_t st, unsigned cmask = ~0u, bool raw_mean = false);
void pb_launch_pass_centered(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                             uint32_t total_blocks, bool weighted, PbStats *d_stats, void *d_scratch,
                             size_t scratch_bytes, cudaStream_t st, unsigned cmask = ~0u);
// Per-bucket ordered sums over bucket-sorted position lists.
//   LQ (local.c:102-146): out[seg][b] = {size (as double bits of u64), sum c0*w, sum c1*w, sum c2*w}
//   GQ (cells.c:53-116):  out[b] = {sum c0, c1, c2, sum |c|^2, sums c_r*c_s (r<=s: 00,01,11,02,12,22)}
void pb_launch_bucket_chains_lq(const double *d_aos, const PbSeg *d_segs, int nseg, bool weighted,
                                const uint32_t *d_ord, const uint32_t *d_class_start,
                                double *d_out /* nseg x 512 x 4 */, cudaStream_t st);
void pb_chain_set_gq_cta(bool on); // test knob: one CTA (default) or one warp per GQ bucket
void pb_launch_bucket_chains_gq(const double *d_aos, const uint32_t *d_ord,
                                const uint32_t *d_class_start, double *d_out /* 512 x 10 */,
                                cudaStream_t st);

// ---- data-parallel kernels, pb_parallel.cu -------------------------------------------
void pb_launch_dots_minmax(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                           const double *d_axes /* nseg x 3 */, PbSplit *d_split, int sm_count,
                           cudaStream_t st);
void pb_launch_buckets(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                       const double *d_axes, PbSplit *d_split, uint16_t *d_bucket, double *d_aos, int sm_count,
                       cudaStream_t st);
// Certified route (pb_certify.cu): bucket ids as pb_launch_buckets plus per-bucket sums in any order (no sort),
// then the proof that their argmax is the reference's; PbSplit::pad says whether it succeeded.
void pb_launch_buckets_hist(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n, bool weighted,
                            const double *d_axes, PbSplit *d_split, uint16_t *d_bucket, PbHist *d_hist, int sm_count,
                            cudaStream_t st);
void pb_launch_split_certify(const PbHist *d_hist, int nseg, bool weighted, PbSplit *d_split, int distrust, cudaStream_t st);
void pb_certify_counts(unsigned long long out[4], bool reset); // {certified, refused, -, -} since the last reset
void pb_launch_split_select(const double *d_bucket_sums, const uint32_t *d_class_start, int nseg,
                            PbSplit *d_split, cudaStream_t st);

// Stable multi-class ranking / scatter.  Class of a position is a function of its bucket
// id; elements keep their relative order inside each class (the reference builds every
// index list by an ascending scan: local.c:216-243, global.c:300-377).
enum { PB_CLS_BUCKET = 0, PB_CLS_SPLIT = 1, PB_CLS_LUT = 2 };
size_t pb_scatter_tiles(uint32_t n);
// d_tile_hist: pb_scatter_table_words(total_tiles, nseg, nclass) u32, packed by PbSeg::tbase (total_tiles
// = capacity of the packed tile table, >= sum of the segments' tiles); d_class_start: nseg * (nclass + 1) u32
size_t pb_scatter_table_words(size_t total_tiles, int nseg, int nclass);
size_t pb_scatter_chunk_offset(size_t total_tiles, int nclass);
void pb_launch_class_rank(int cls_mode, int nclass, const PbSeg *d_segs, int nseg, uint32_t max_n,
                          size_t total_tiles, const uint16_t *d_bucket, const PbSplit *d_split,
                          const uint8_t *d_lut, uint32_t *d_tile_hist, uint32_t *d_class_start, cudaStream_t st);
// ord[dst] = source position (the bucket sort feeding the ordered per-bucket sums)
void pb_launch_scatter_ord(int cls_mode, int nclass, const PbSeg *d_segs, int nseg, uint32_t max_n,
                           const uint16_t *d_bucket, const PbSplit *d_split, const uint8_t *d_lut,
                           const uint32_t *d_tile_hist, const uint32_t *d_class_start, uint32_t *d_ord,
                           cudaStream_t st);
// sorted[dst] = (c0, c1, c2, w) of the source pixel: the bucket-sorted interleaved copy the per-bucket sums read
// sequentially (d_ord == nullptr in the pb_launch_bucket_chains_* calls below)
void pb_scatter_set_cta(bool on); // test knob
void pb_launch_scatter_sorted(int cls_mode, int nclass, const PbPlanes src[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                              const uint16_t *d_bucket, const PbSplit *d_split, const uint8_t *d_lut,
                              const uint32_t *d_tile_hist, const uint32_t *d_class_start, double *d_sorted, cudaStream_t st);
// move the payload (planes + weight + original index); segment s reads src[s.buf], writes dst[s.buf]
void pb_launch_scatter_payload(int cls_mode, int nclass, const PbPlanes src[2], const PbPlanes dst[2],
                               bool src_is_identity, const PbSeg *d_segs, int nseg, uint32_t max_n,
                               const uint16_t *d_bucket, const PbSplit *d_split, const uint8_t *d_lut,
                               const uint32_t *d_tile_hist, const uint32_t *d_class_start, cudaStream_t st);
// child descriptors {lo, nleft, other buf}, {lo + nleft, n - nleft, other buf} from PbSplit
void pb_launch_make_children(const PbSeg *d_segs, int nseg, const PbSplit *d_split,
                             PbSeg *d_children /* 2 x nseg */, cudaStream_t st);

// Nearest palette entry + cluster labels.
void pb_launch_nearest(const double *const planes[3], size_t n, const double *d_palette_rm, int K,
                       unsigned long long *d_map, int sm_count, cudaStream_t st);
// exact 1-NN through per-cell candidate lists (pb_nngrid.cu): same map as pb_launch_nearest
size_t pb_nngrid_scratch_bytes(int K);
// scratch layout: PbGridHdr (256 B) | u16 count[4096] | u16 list[4096][K].  `scale` (nullptr = 1) maps pixel space to
// query space, `expand` widens the pixels' box by that fraction of its range on every side.
void pb_launch_nngrid_build(const double *const planes[3], size_t n, const double *d_palette_rm, int K, void *d_scratch,
                            int sm_count, cudaStream_t st, const double *scale = nullptr, double expand = 0.0);
void pb_launch_nearest_grid(const double *const planes[3], size_t n, const double *d_palette_rm, int K, const void *d_scratch,
                            unsigned long long *d_map, int sm_count, cudaStream_t st);
void pb_launch_labels(const PbPlanes bufs[2], const PbSeg *d_segs, int nseg, uint32_t max_n,
                      uint32_t *d_labels, int sm_count, cudaStream_t st);
