// cosmomap2_b200 -- single-TOD-pass y = P^T F P x for the subscan offset filter (sm_100a).
//
// F = I - Pi, Pi = per-subscan mean over the unflagged samples (FilterLO.mult,
// interfaces/linearoperators.py:129-168), so
//     P^T F P x = P^T P x  -  sum_k mu_k u_k ,   u_k = P^T 1_k ,  mu_k = (u_k . x) / n_k
// where 1_k is the indicator of the unflagged samples of subscan k.  u_k only depends on the
// pointing: it is stored once, run-compressed (a scan crosses a pixel in a run of consecutive
// samples): per run the pixel, the sample count n and sum(cos), sum(sin) -- 28 B per RUN instead of
// 20 B per SAMPLE.  An A-matvec is then
//     k_seg_mean         : CTA per subscan over the run table: mu_k = sum_runs (n I_p + C Q_p + S U_p) / sum n
//     k_amatvec_filter_mu: ONE pass over the TOD (tod_pass.cu): y += P^T (P x - mu_seg(t)) on the
//                          unflagged samples inside subscans -- the fused white kernel with a
//                          per-chunk segment lookup instead of a block-weight lookup.
// (A first version scattered the correction -mu_k u_k from the run table instead; its second RED
// stream cost 0.17 ms at configs[1] size, the mean pre-pass costs a fraction of that.)
// The two-pass kernel in tod_pass.cu (k_amatvec_filter) is kept as the no-extra-memory variant.
#include "cm2_common.cuh"

namespace cm2 {

constexpr int FB = 256;

// flags[t] = 1 where a run of equal pixel starts among the unflagged samples of a segment;
// pix_masked[t] = pix[t] inside segments, -1 elsewhere (pre-filled with -1 by the caller)
__global__ void __launch_bounds__(FB) k_runs_mark(const int32_t *__restrict__ pix, const int64_t *__restrict__ seg_start,
                                                  const int64_t *__restrict__ seg_end, int64_t nseg,
                                                  int32_t *__restrict__ flags, int32_t *__restrict__ pix_masked) {
    for (int64_t k = blockIdx.x; k < nseg; k += gridDim.x) {
        const int64_t a = seg_start[k], b = seg_end[k];
        for (int64_t t = a + threadIdx.x; t < b; t += FB) {
            const int32_t p = pix[t];
            const int32_t prev = t > a ? pix[t - 1] : -1;
            flags[t] = (p >= 0 && (t == a || prev != p)) ? 1 : 0;
            pix_masked[t] = p;
        }
    }
}

// one thread per run start: walk the run, accumulate (n, sum cos, sum sin); per segment the first
// run index and the number of runs
__global__ void __launch_bounds__(FB) k_runs_fill(const int32_t *__restrict__ pix, const double *__restrict__ cs,
                                                  const double *__restrict__ sn, int pol,
                                                  const int64_t *__restrict__ seg_start, const int64_t *__restrict__ seg_end,
                                                  int64_t nseg, const int32_t *__restrict__ runidx,
                                                  int32_t *__restrict__ run_pix, double *__restrict__ run_mom,
                                                  int64_t *__restrict__ seg_first, int32_t *__restrict__ seg_nruns) {
    __shared__ int s_first, s_count;
    for (int64_t k = blockIdx.x; k < nseg; k += gridDim.x) {
        const int64_t a = seg_start[k], b = seg_end[k];
        if (threadIdx.x == 0) { s_first = 0x7fffffff; s_count = 0; }
        __syncthreads();
        for (int64_t t = a + threadIdx.x; t < b; t += FB) {
            const int32_t r = runidx[t];
            if (r < 0) continue;
            const int32_t p = pix[t];
            double n = 0.0, c = 0.0, s = 0.0;
            for (int64_t u = t; u < b && pix[u] == p; ++u) {
                n += 1.0;
                if (pol > 1) { c += cs[u]; s += sn[u]; }
            }
            run_pix[r] = p;
            run_mom[3 * (int64_t)r] = n;
            run_mom[3 * (int64_t)r + 1] = c;
            run_mom[3 * (int64_t)r + 2] = s;
            atomicMin(&s_first, r);
            atomicAdd(&s_count, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            seg_first[k] = s_count > 0 ? s_first : 0;
            seg_nruns[k] = s_count;
        }
        __syncthreads();
    }
}

template <int POL>
__global__ void __launch_bounds__(FB) k_seg_mean(const int32_t *__restrict__ run_pix, const double *__restrict__ run_mom,
                                                 const int64_t *__restrict__ seg_first, const int32_t *__restrict__ seg_nruns,
                                                 int64_t nseg, const double *__restrict__ x, double *__restrict__ mu) {
    __shared__ double red[32];
    for (int64_t k = blockIdx.x; k < nseg; k += gridDim.x) {
        const int64_t f = seg_first[k];
        const int nr = seg_nruns[k];
        double dsum = 0.0, cnt = 0.0;
        for (int i = threadIdx.x; i < nr; i += FB) {
            const int64_t r = f + i;
            const int32_t p = __ldcs(run_pix + r);
            const double n = __ldcs(run_mom + 3 * r), c = __ldcs(run_mom + 3 * r + 1), s = __ldcs(run_mom + 3 * r + 2);
            const double *xp = x + (int64_t)POL * p;
            if constexpr (POL == 1) dsum = fma(n, __ldg(xp), dsum);
            else if constexpr (POL == 2) dsum = fma(s, __ldg(xp + 1), fma(c, __ldg(xp), dsum));
            else dsum = fma(s, __ldg(xp + 2), fma(c, __ldg(xp + 1), fma(n, __ldg(xp), dsum)));
            cnt += n;
        }
        const double ts = block_sum(dsum, red);
        const double tc = block_sum(cnt, red);
        if (threadIdx.x == 0) mu[k] = tc > 0.0 ? ts / tc : 0.0;
    }
}

static int fgrid(int64_t n) {
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(n < 1 ? 1 : (n < cap ? n : cap));
}

}  // namespace cm2

using namespace cm2;

extern "C" int cm2_filter_runs_mark(const int32_t *pix, const int64_t *seg_start, const int64_t *seg_end, int64_t nseg,
                                    int64_t nt, int32_t *flags, int32_t *pix_masked, cm2_stream_t stream) {
    CM2_REQUIRE(nt >= 0 && nseg >= 0, "bad sizes");
    cudaStream_t st = as_stream(stream);
    if (nt > 0) {
        CM2_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t) * (size_t)nt, st));
        CM2_CUDA(cudaMemsetAsync(pix_masked, 0xff, sizeof(int32_t) * (size_t)nt, st));   // -1 everywhere
    }
    if (nt == 0 || nseg == 0) return CM2_OK;
    k_runs_mark<<<fgrid(nseg), FB, 0, st>>>(pix, seg_start, seg_end, nseg, flags, pix_masked);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_filter_runs_fill(const int32_t *pix, const double *c, const double *s, int pol,
                                    const int64_t *seg_start, const int64_t *seg_end, int64_t nseg,
                                    const int32_t *runidx, int32_t *run_pix, double *run_mom, int64_t *seg_first,
                                    int32_t *seg_nruns, cm2_stream_t stream) {
    CM2_REQUIRE(nseg >= 0 && pol >= 1 && pol <= 3, "bad sizes");
    if (nseg == 0) return CM2_OK;
    k_runs_fill<<<fgrid(nseg), FB, 0, as_stream(stream)>>>(pix, c, s, pol, seg_start, seg_end, nseg, runidx, run_pix, run_mom,
                                                         seg_first, seg_nruns);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_filter_seg_mean(const int32_t *run_pix, const double *run_mom, const int64_t *seg_first,
                                   const int32_t *seg_nruns, int64_t nseg, int pol, const double *x, double *mu,
                                   cm2_stream_t stream) {
    CM2_REQUIRE(nseg >= 0 && pol >= 1 && pol <= 3, "bad sizes");
    if (nseg == 0) return CM2_OK;
    cudaStream_t st = as_stream(stream);
    if (pol == 1) k_seg_mean<1><<<fgrid(nseg), FB, 0, st>>>(run_pix, run_mom, seg_first, seg_nruns, nseg, x, mu);
    else if (pol == 2) k_seg_mean<2><<<fgrid(nseg), FB, 0, st>>>(run_pix, run_mom, seg_first, seg_nruns, nseg, x, mu);
    else k_seg_mean<3><<<fgrid(nseg), FB, 0, st>>>(run_pix, run_mom, seg_first, seg_nruns, nseg, x, mu);
    CM2_LAUNCHED();
    return CM2_OK;
}
