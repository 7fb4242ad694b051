// cosmomap2_b200 -- SURVEY 8(f) rows either side of the solve (sm_100a):
//
//   cm2_filter_poly_apply     subscan filter, one CTA per subscan, the subscan staged in shared memory
//                             poly_order = 0 : offset removal        (FilterLO.mult       linearoperators.py:129-168)
//                             poly_order > 0 : Legendre polynomials  (FilterLO.polyfilter linearoperators.py:170-204)
//   cm2_ground_filter_apply   v - G (G^T G)^-1 G^T v over ground bins (GroundFilterLO     linearoperators.py:24-61)
//   cm2_reorganize_map        interleaved cut-sky solution -> full-sky HEALPix arrays
//                                                                    (reorganize_map      healpy_functions.py:50-105)
//
// The subscan filter reads d and pix ONCE (12 B/sample) and writes out once (8 B/sample): the samples
// of a subscan wait in shared memory between the reduction pass and the subtraction pass, and the
// CTA of subscan k also zero-fills the gap in front of it, so no memset of the output is needed.
#include <climits>

#include "cm2_common.cuh"

namespace cm2 {

constexpr int FB = 256;          // threads per CTA
constexpr int FW = FB / 32;      // warps per CTA
constexpr int FP_MAXNK = 8;      // poly_order <= 7

// NK = poly_order + 1.  OFFSET (NK == 1): the reference's poly_order = 0 path, where the mean over
// the unflagged (pix != -1) samples is subtracted from EVERY sample of the subscan (:165).
// Otherwise the Legendre path: unflagged = pix >= 0 (:174); a subscan with <= poly_order unflagged
// samples is skipped (:185-187); without flags p = sum_k (b_k . d) b_k with b_k = L_k/||L_k|| on the
// full grid (:196-200, NOT an exact projector: the sampled L_k are not orthogonal); with flags the
// basis is re-orthonormalised by QR on the unflagged rows (:190-194), i.e. p is the least-squares
// polynomial of degree <= poly_order -- computed here from the Gram matrix in the Legendre basis of
// the interval spanned by the unflagged samples (same span, well conditioned); flagged samples -> 0.
// fill != 0: every output sample is written by this kernel (segments sorted, non-overlapping).
// fill == 0: the caller zero-filled `out`; only what the reference assigns is written.
template <int NK, bool OFFSET>
__global__ void __launch_bounds__(FB) k_filter_poly(const int32_t *__restrict__ pix, const int64_t *__restrict__ seg_start,
                                                    const int64_t *__restrict__ seg_end, int64_t nseg,
                                                    const double *__restrict__ d, double *__restrict__ out, int64_t nt,
                                                    int cap, int fill) {
    constexpr int NG = OFFSET ? 0 : NK * (NK + 1) / 2;
    constexpr int NR = NK + NG;
    extern __shared__ double sm[];
    double *sd = sm;                                     // cap staged samples
    uint8_t *sf = reinterpret_cast<uint8_t *>(sm + cap);   // cap flags (1 = unflagged)
    __shared__ double red[FW * NR];
    __shared__ double tot[NR];
    __shared__ double coef[NK];
    __shared__ int s_cnt, s_jmin, s_jmax, s_skip, s_refine;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int64_t k = blockIdx.x; k < nseg; k += gridDim.x) {
        const int64_t a = seg_start[k], b = seg_end[k];
        const int64_t len = b - a;
        if (fill) {          // gap in front of this subscan (and behind the last one)
            const int64_t g0 = k == 0 ? 0 : seg_end[k - 1];
            for (int64_t t = g0 + tid; t < a; t += FB) __stcs(out + t, 0.0);
            if (k == nseg - 1)
                for (int64_t t = b + tid; t < nt; t += FB) __stcs(out + t, 0.0);
        }
        if (tid == 0) { s_cnt = 0; s_jmin = INT_MAX; s_jmax = -1; }
        __syncthreads();
        // ---- pass 1: stage, count, (offset path) sum ------------------------------------------
        int cnt = 0, jmin = INT_MAX, jmax = -1;
        double sum = 0.0;
        for (int64_t j = tid; j < len; j += FB) {
            const int p = __ldcs(pix + a + j);
            const double v = __ldcs(d + a + j);
            const bool f = OFFSET ? (p != -1) : (p >= 0);
            if (j < cap) { sd[j] = v; sf[j] = f; }
            if (f) {
                ++cnt;
                if (jmin == INT_MAX) jmin = (int)j;
                jmax = (int)j;
                if (OFFSET) sum += v;
            }
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        jmin = __reduce_min_sync(0xffffffffu, jmin);
        jmax = __reduce_max_sync(0xffffffffu, jmax);
        if (lane == 0 && cnt > 0) { atomicAdd(&s_cnt, cnt); atomicMin(&s_jmin, jmin); atomicMax(&s_jmax, jmax); }
        if constexpr (OFFSET) {
            const double tsum = block_sum(sum, red);     // two barriers inside: s_cnt is complete after it
            if (tid == 0) {
                const double mean = tsum / (double)s_cnt;     // 0/0 = NaN like the reference (:154, 163)
                s_skip = (s_cnt == 0) || isinf(mean) || isnan(mean);
                coef[0] = mean;
            }
            __syncthreads();
        } else {
            __syncthreads();
            const int n = s_cnt;
            const bool skip = n <= NK - 1;
            if (!skip) {
                const bool full = (int64_t)n == len;
                const int j0 = full ? 0 : s_jmin;
                const int j1 = full ? (int)(len - 1) : s_jmax;
                const double step = 2.0 / (double)(j1 - j0);
                // ---- pass 2: S_k = sum L_k d, G_kl = sum L_k L_l over the unflagged samples -----
                double acc[NR];
#pragma unroll
                for (int i = 0; i < NR; ++i) acc[i] = 0.0;
                for (int64_t j = tid; j < len; j += FB) {
                    const bool f = j < cap ? (sf[j] != 0) : (pix[a + j] >= 0);
                    if (!f) continue;
                    const double v = j < cap ? sd[j] : d[a + j];
                    double L[NK];
                    legendre<NK>(fma((double)((int)j - j0), step, -1.0), L);
                    int q = NK;
#pragma unroll
                    for (int r = 0; r < NK; ++r) {
                        acc[r] = fma(L[r], v, acc[r]);
#pragma unroll
                        for (int c = r; c < NK; ++c) { acc[q] = fma(L[r], L[c], acc[q]); ++q; }
                    }
                }
                block_sum_n<NR, FW>(acc, red, tot);
                if (tid == 0) {
                    s_refine = 0;
                    if (full) {
                        int q = NK;
                        for (int r = 0; r < NK; ++r) {
                            coef[r] = tot[q] > 0.0 ? tot[r] / tot[q] : 0.0;   // (b_k . d) / ||L_k||^2
                            q += NK - r;
                        }
                    } else {
                        double c[NK];
                        s_refine = filter_refine_steps(gram_solve<NK>(tot + NK, tot, c));
                        for (int r = 0; r < NK; ++r) coef[r] = c[r];
                    }
                }
                __syncthreads();
                const int nref = s_refine;                   // block-uniform, 0 unless ill-conditioned
                for (int it = 0; it < nref; ++it) {
                    double c[NK], racc[NK];
#pragma unroll
                    for (int r = 0; r < NK; ++r) { c[r] = coef[r]; racc[r] = 0.0; }
                    for (int64_t j = tid; j < len; j += FB) {
                        const bool f = j < cap ? (sf[j] != 0) : (pix[a + j] >= 0);
                        if (!f) continue;
                        double L[NK];
                        legendre<NK>(fma((double)((int)j - j0), step, -1.0), L);
                        double res = j < cap ? sd[j] : d[a + j];
#pragma unroll
                        for (int r = 0; r < NK; ++r) res = fma(-c[r], L[r], res);
#pragma unroll
                        for (int r = 0; r < NK; ++r) racc[r] = fma(L[r], res, racc[r]);
                    }
                    block_sum_n<NK, FW>(racc, red, tot);      // tot[0..NK) = L^T (d - L c); the Gram matrix stays
                    if (tid == 0) {
                        double dc[NK];
                        gram_solve<NK>(tot + NK, tot, dc);
                        for (int r = 0; r < NK; ++r) coef[r] += dc[r];
                    }
                    __syncthreads();
                }
            }
            if (tid == 0) s_skip = skip;
            __syncthreads();
        }
        // ---- pass 3: subtract and write --------------------------------------------------------
        const bool skip = s_skip != 0;
        if (OFFSET) {
            const double mu = coef[0];
            if (!skip) {
                for (int64_t j = tid; j < len; j += FB) __stcs(out + a + j, (j < cap ? sd[j] : d[a + j]) - mu);
            } else if (fill) {
                for (int64_t j = tid; j < len; j += FB) __stcs(out + a + j, 0.0);
            }
        } else {
            if (!skip) {
                const bool full = (int64_t)s_cnt == len;
                const int j0 = full ? 0 : s_jmin;
                const int j1 = full ? (int)(len - 1) : s_jmax;
                const double step = 2.0 / (double)(j1 - j0);
                double c[NK];
#pragma unroll
                for (int r = 0; r < NK; ++r) c[r] = coef[r];
                for (int64_t j = tid; j < len; j += FB) {
                    const bool f = j < cap ? (sf[j] != 0) : (pix[a + j] >= 0);
                    if (f) {
                        const double v = j < cap ? sd[j] : d[a + j];
                        double L[NK];
                        legendre<NK>(fma((double)((int)j - j0), step, -1.0), L);
                        double p = 0.0;
#pragma unroll
                        for (int r = 0; r < NK; ++r) p = fma(c[r], L[r], p);
                        __stcs(out + a + j, v - p);
                    } else if (fill) {
                        __stcs(out + a + j, 0.0);
                    }
                }
            } else if (fill) {
                for (int64_t j = tid; j < len; j += FB) __stcs(out + a + j, 0.0);
            }
        }
        __syncthreads();
    }
}

template <int NK, bool OFFSET>
static int launch_filter_poly(const int32_t *pix, const int64_t *seg_start, const int64_t *seg_end, int64_t nseg,
                              const double *d, double *out, int64_t nt, int cap, int fill, cudaStream_t st) {
    const size_t smem = (size_t)cap * 9;
    auto kern = k_filter_poly<NK, OFFSET>;
    CM2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<persistent_grid(kern, FB, smem, nseg), FB, smem, st>>>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill);
    CM2_LAUNCHED();
    return CM2_OK;
}

// ---- ground-template filter: out = v - bins[g] / hits[g] ---------------------------------------
__global__ void __launch_bounds__(FB) k_ground_sub(const int32_t *__restrict__ g, const double *__restrict__ bins,
                                                   const int64_t *__restrict__ hits, const double *__restrict__ v,
                                                   double *__restrict__ out, int64_t nt) {
    const int64_t nchunk = (nt + 3) / 4;
    for (int64_t ch = (int64_t)blockIdx.x * FB + threadIdx.x; ch < nchunk; ch += (int64_t)gridDim.x * FB) {
        const int64_t t0 = ch * 4;
        if (t0 + 4 <= nt) {
            const int4 gi = __ldcs(reinterpret_cast<const int4 *>(g + t0));
            D4 x = ld_stream_d4(v + t0);
            const int gg[4] = {gi.x, gi.y, gi.z, gi.w};
            int prev = -1;
            double m = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (gg[j] >= 0) {
                    if (gg[j] != prev) {          // ground bins change slowly along a scan: one lookup per run
                        const int64_t h = __ldg(hits + gg[j]);
                        m = h > 0 ? __ldg(bins + gg[j]) / (double)h : 0.0;   // 1x1 M_BD: x / counts where counts > 0
                        prev = gg[j];
                    }
                    x.v[j] -= m;
                }
            }
            st_stream_d4(out + t0, x);
        } else {
            for (int64_t t = t0; t < nt; ++t) {
                const int b = g[t];
                double m = 0.0;
                if (b >= 0) { const int64_t h = hits[b]; m = h > 0 ? bins[b] / (double)h : 0.0; }
                out[t] = v[t] - m;
            }
        }
    }
}

// ---- map output: out[k][obspix[p]] = map[pol p + k], zero elsewhere ----------------------------------
__global__ void __launch_bounds__(FB) k_reorganize(const double *__restrict__ map, const int64_t *__restrict__ obspix,
                                                   int64_t npix, int pol, int64_t hnpix, double *__restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * FB + threadIdx.x; p < npix; p += (int64_t)gridDim.x * FB) {
        const int64_t o = obspix[p];
        for (int k = 0; k < pol; ++k) out[(int64_t)k * hnpix + o] = map[(int64_t)pol * p + k];
    }
}

static int grid_fp(int64_t blocks, int per_sm = 8) {
    int64_t capb = (int64_t)sm_count() * per_sm;
    if (blocks > capb) blocks = capb;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace cm2

using namespace cm2;

extern "C" int cm2_filter_poly_max_order(void) { return FP_MAXNK - 1; }

extern "C" int cm2_filter_poly_apply(const int32_t *pix, const int64_t *seg_start, const int64_t *seg_end, int64_t nseg,
                                     int64_t max_seg_len, int poly_order, int sorted, const double *d, double *out,
                                     int64_t nt, cm2_stream_t stream) {
    CM2_REQUIRE(nt >= 0 && nseg >= 0 && max_seg_len >= 0, "bad sizes");
    CM2_REQUIRE(d != out, "in-place filtering is not supported");
    CM2_REQUIRE(max_seg_len < INT_MAX, "subscans longer than 2^31 samples are not supported");
    if (poly_order < 0 || poly_order >= FP_MAXNK)
        return set_error(CM2_ERR_UNSUPPORTED, "poly_order=%d: orders 0..%d are supported", poly_order, FP_MAXNK - 1);
    cudaStream_t st = as_stream(stream);
    const int fill = sorted && nseg > 0;
    if (nt > 0 && !fill) CM2_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)nt, st));
    if (nt == 0 || nseg == 0) return CM2_OK;
    // shared-memory window: the longest subscan, up to 24 000 samples (216 kB); longer subscans
    // re-read their tail from global memory (L2)
    int64_t cap64 = (max_seg_len + 15) / 16 * 16;
    if (cap64 > 24000) cap64 = 24000;
    if (cap64 < 16) cap64 = 16;
    const int cap = (int)cap64;
    switch (poly_order) {
        case 0: return launch_filter_poly<1, true>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill, st);
        case 1: return launch_filter_poly<2, false>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill, st);
        case 2: return launch_filter_poly<3, false>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill, st);
        case 3: return launch_filter_poly<4, false>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill, st);
        case 4: return launch_filter_poly<5, false>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill, st);
        case 5: return launch_filter_poly<6, false>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill, st);
        case 6: return launch_filter_poly<7, false>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill, st);
        default: return launch_filter_poly<8, false>(pix, seg_start, seg_end, nseg, d, out, nt, cap, fill, st);
    }
}

extern "C" int cm2_ground_filter_apply(const int32_t *ground, int64_t nt, int64_t nbins, const int64_t *hits,
                                       const double *v, double *bins, double *out, cm2_stream_t stream) {
    CM2_REQUIRE(nt >= 0 && nbins >= 0, "bad sizes");
    CM2_REQUIRE(aligned(ground, 32) && aligned(v, 32) && aligned(out, 32), "TOD vectors must be 32-byte aligned");
    // bins = G^T v: the pol = 1 scatter-add with run aggregation (ground bins change slowly along a scan)
    int rc = cm2_pointing_apply_t(ground, nullptr, nullptr, nt, 1, v, bins, nbins, stream);
    if (rc) return rc;
    if (nt == 0) return CM2_OK;
    k_ground_sub<<<grid_fp(((nt + 3) / 4 + FB - 1) / FB), FB, 0, as_stream(stream)>>>(ground, bins, hits, v, out, nt);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_reorganize_map(const double *map, const int64_t *obspix, int64_t npix, int pol, int64_t healpix_npix,
                                  double *out, cm2_stream_t stream) {
    CM2_REQUIRE(npix >= 0 && healpix_npix >= 0, "bad sizes");
    CM2_REQUIRE(pol >= 1 && pol <= 3, "pol must be 1, 2 or 3");
    cudaStream_t st = as_stream(stream);
    if (healpix_npix > 0) CM2_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)healpix_npix * pol, st));
    if (npix == 0 || healpix_npix == 0) return CM2_OK;
    k_reorganize<<<grid_fp((npix + FB - 1) / FB), FB, 0, st>>>(map, obspix, npix, pol, healpix_npix, out);
    CM2_LAUNCHED();
    return CM2_OK;
}
