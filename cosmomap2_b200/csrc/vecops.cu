// cosmomap2_b200 -- PCG vector work on the device (sm_100a).
//
// Restates the recurrence of scipy.sparse.linalg.cg (scipy/_isolve/iterative.py:405-431), which the
// reference calls at src/test_BD_precond_onto_real_data.py:47 and
// src/test_M2_precond_onto_real_data.py:117.  All scalars live in a 16-double device workspace:
//   scal[0]=rho  [1]=rho_prev  [2]=p.q  [3]=|r|^2  [4]=alpha  [5]=beta  [6]=atol
//   scal[7]=done flag (||r|| < atol, SciPy's exit test)  [8]=iterations completed  [9..15] spare
// so an iteration needs no host round trip, and every update kernel is a no-op once `done` is set:
// the host may launch iterations ahead of reading the flag and x still freezes at exactly the
// iteration SciPy would have returned.
//
// Reductions are deterministic: fixed per-thread order, fixed shuffle tree, per-CTA partials summed
// in CTA order by the last CTA to finish (ticket).  The partial/ticket scratch is one device-global
// area: call these entry points from one stream at a time.
#include <cooperative_groups.h>

#include "cm2_common.cuh"

namespace cg = cooperative_groups;

namespace cm2 {

constexpr int VB = 256;
constexpr int MAXP = 2048;  // max CTAs of a reduction kernel

__device__ double g_part[3][MAXP];   // rows 0,1: reduction partials; row 2: p.q partials of k_bd_iter
__device__ unsigned int g_ticket;

// CTA-level: write up to two partials, last CTA reduces all partials in order; returns true in
// thread 0 of the last CTA with the totals in tot[0..NR)
template <int NR>
__device__ __forceinline__ bool finish_reduce(const double (&v)[NR], double *red, double (&tot)[NR]) {
    __shared__ bool last;
    double t[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) t[k] = block_sum(v[k], red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) g_part[k][blockIdx.x] = t[k];
        __threadfence();
        const unsigned int tk = atomicAdd(&g_ticket, 1u);
        last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += VB) s += ((volatile double *)g_part[k])[i];
        tot[k] = block_sum(s, red);
    }
    if (threadIdx.x == 0) {
        g_ticket = 0;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(VB) k_dot(const double *__restrict__ a, const double *__restrict__ b, int64_t n,
                                            double *__restrict__ out) {
    __shared__ double red[32];
    double s[1] = {0.0}, tot[1];
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) s[0] = fma(a[i], b[i], s[0]);
    if (finish_reduce<1>(s, red, tot)) *out = tot[0];
}

__global__ void __launch_bounds__(VB) k_axpby(double alpha, const double *__restrict__ x, double beta, double *__restrict__ y,
                                              int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        const double yi = beta == 0.0 ? 0.0 : beta * y[i];
        y[i] = fma(alpha, x[i], yi);
    }
}

// ---- generic-M path ------------------------------------------------------------------------
// reset: |r|^2, atol, done flag, iteration counter
__global__ void __launch_bounds__(VB) k_reset(const double *__restrict__ r, int64_t n, double *__restrict__ scal, double atol) {
    __shared__ double red[32];
    double s[1] = {0.0}, tot[1];
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) s[0] = fma(r[i], r[i], s[0]);
    if (finish_reduce<1>(s, red, tot)) {
        scal[0] = 0.0; scal[1] = 0.0; scal[2] = 0.0; scal[4] = 0.0; scal[5] = 0.0;
        scal[3] = tot[0];
        scal[6] = atol;
        scal[7] = (sqrt(tot[0]) < atol) ? 1.0 : 0.0;
        scal[8] = 0.0;
    }
}

// rho = r.z ; beta = rho/rho_prev (0 on the first iteration)
__global__ void __launch_bounds__(VB) k_rho(const double *__restrict__ r, const double *__restrict__ z, int64_t n,
                                            double *__restrict__ scal) {
    __shared__ double red[32];
    if (scal[7] != 0.0) return;
    double s[1] = {0.0}, tot[1];
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) s[0] = fma(r[i], z[i], s[0]);
    if (finish_reduce<1>(s, red, tot)) {
        scal[0] = tot[0];
        scal[5] = (scal[8] == 0.0) ? 0.0 : tot[0] / scal[1];
    }
}

// p = z + beta p  (p = z on the first iteration: p may be uninitialised)
__global__ void __launch_bounds__(VB) k_update_p(const double *__restrict__ z, double *__restrict__ p, int64_t n,
                                                 const double *__restrict__ scal) {
    if (scal[7] != 0.0) return;
    const bool first = scal[8] == 0.0;
    const double beta = scal[5];
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB)
        p[i] = first ? z[i] : fma(beta, p[i], z[i]);
}

__global__ void __launch_bounds__(VB) k_pq(const double *__restrict__ p, const double *__restrict__ q, int64_t n,
                                           double *__restrict__ scal) {
    __shared__ double red[32];
    if (scal[7] != 0.0) return;
    double s[1] = {0.0}, tot[1];
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) s[0] = fma(p[i], q[i], s[0]);
    if (finish_reduce<1>(s, red, tot)) {
        scal[2] = tot[0];
        scal[4] = scal[0] / tot[0];
    }
}

__global__ void __launch_bounds__(VB) k_update_xr(const double *__restrict__ p, const double *__restrict__ q,
                                                  double *__restrict__ x, double *__restrict__ r, int64_t n,
                                                  double *__restrict__ scal) {
    __shared__ double red[32];
    if (scal[7] != 0.0) return;
    const double alpha = scal[4];
    double s[1] = {0.0}, tot[1];
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        r[i] = ri;
        s[0] = fma(ri, ri, s[0]);
    }
    if (finish_reduce<1>(s, red, tot)) {
        scal[3] = tot[0];
        scal[1] = scal[0];
        scal[8] += 1.0;
        scal[7] = (sqrt(tot[0]) < scal[6]) ? 1.0 : 0.0;
    }
}

// ---- M = M_BD path: the preconditioner apply is pixel-local, so z = M r and rho = r.z ride in
// the same kernel that updates r (one pass over the pixel-domain vectors per iteration) ---------
// start: z = M r, rho = r.z, |r|^2, flags.  With b != nullptr it also performs the x0 = 0 start of
// the solve in the same pass: r = b, x = 0, and folds SciPy's atol = max(atol, rtol ||b||) in (rtol > 0).
template <int POL>
__global__ void __launch_bounds__(VB) k_bd_reset(const double *__restrict__ inv, int64_t npix, double *__restrict__ r,
                                                 double *__restrict__ z, double *__restrict__ scal, double atol, double rtol,
                                                 const double *__restrict__ b, double *__restrict__ x,
                                                 double *__restrict__ p0) {
    __shared__ double red[32];
    double s[2] = {0.0, 0.0}, tot[2];
    for (int64_t j = (int64_t)blockIdx.x * VB + threadIdx.x; j < npix; j += (int64_t)gridDim.x * VB) {
        double rv[POL], zv[POL];
        if (b != nullptr) {
#pragma unroll
            for (int k = 0; k < POL; ++k) {
                rv[k] = b[POL * j + k];
                r[POL * j + k] = rv[k];
                x[POL * j + k] = 0.0;
            }
        } else {
#pragma unroll
            for (int k = 0; k < POL; ++k) rv[k] = r[POL * j + k];
        }
        bd_z<POL>(inv, j, rv, zv);
#pragma unroll
        for (int k = 0; k < POL; ++k) {
            z[POL * j + k] = zv[k];
            if (p0 != nullptr) p0[POL * j + k] = zv[k];     // first search direction p = z
            s[0] = fma(rv[k], zv[k], s[0]);
            s[1] = fma(rv[k], rv[k], s[1]);
        }
    }
    if (finish_reduce<2>(s, red, tot)) {
        const double atol_eff = (b != nullptr && rtol > 0.0) ? fmax(atol, rtol * sqrt(tot[1])) : atol;
        scal[0] = tot[0]; scal[1] = 0.0; scal[2] = 0.0; scal[4] = 0.0; scal[5] = 0.0;
        scal[3] = tot[1];
        scal[6] = atol_eff;
        scal[7] = (sqrt(tot[1]) < atol_eff || (b != nullptr && tot[1] == 0.0)) ? 1.0 : 0.0;
        scal[8] = 0.0;
        scal[9] = 0.0;                      // the sharded solver's failure word
        scal[10] = tot[1];                  // ||r||^2 at the start (||b||^2 for x0 = 0)
    }
}

// alpha = rho/pq ; x += alpha p ; r -= alpha q ; z = M r ; rho' = r.z ; |r|^2 ; beta' = rho'/rho
template <int POL>
__global__ void __launch_bounds__(VB) k_bd_update(const double *__restrict__ inv, int64_t npix, const double *__restrict__ p,
                                                  const double *__restrict__ q, double *__restrict__ x,
                                                  double *__restrict__ r, double *__restrict__ z, double *__restrict__ scal) {
    __shared__ double red[32];
    if (scal[7] != 0.0) return;
    const double alpha = scal[4];
    double s[2] = {0.0, 0.0}, tot[2];
    for (int64_t j = (int64_t)blockIdx.x * VB + threadIdx.x; j < npix; j += (int64_t)gridDim.x * VB) {
        double rv[POL], zv[POL];
#pragma unroll
        for (int k = 0; k < POL; ++k) {
            const int64_t i = POL * j + k;
            x[i] = fma(alpha, p[i], x[i]);
            rv[k] = fma(-alpha, q[i], r[i]);
            r[i] = rv[k];
        }
        bd_z<POL>(inv, j, rv, zv);
#pragma unroll
        for (int k = 0; k < POL; ++k) {
            z[POL * j + k] = zv[k];
            s[0] = fma(rv[k], zv[k], s[0]);
            s[1] = fma(rv[k], rv[k], s[1]);
        }
    }
    if (finish_reduce<2>(s, red, tot)) {
        const double rho_old = scal[0];
        scal[1] = rho_old;
        scal[0] = tot[0];
        scal[5] = tot[0] / rho_old;
        scal[3] = tot[1];
        scal[8] += 1.0;
        scal[7] = (sqrt(tot[1]) < scal[6]) ? 1.0 : 0.0;
    }
}

// ---- the whole pixel-domain tail of an M_BD iteration in ONE cooperative launch -----------------
//   pq = p.q ; alpha = rho/pq ; x += alpha p ; r -= alpha q ; z = M r ; rho' = r.z ; |r|^2 ;
//   beta' = rho'/rho ; p = z + beta' p   (the NEXT iteration's search direction) ; flags
// Two grid-wide barriers replace two kernel boundaries; reductions stay deterministic (per-CTA
// partials summed in CTA order, by every CTA identically).
__device__ __forceinline__ double sum_partials(const double *part, int nb, double *red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nb; i += VB) s += ((volatile const double *)part)[i];
    __shared__ double bc;
    const double t = block_sum(s, red);
    if (threadIdx.x == 0) bc = t;
    __syncthreads();
    const double out = bc;
    __syncthreads();
    return out;
}

template <int POL>
__global__ void __launch_bounds__(VB) k_bd_iter(const double *__restrict__ inv, int64_t npix, double *__restrict__ p,
                                                const double *__restrict__ q, double *__restrict__ x, double *__restrict__ r,
                                                double *__restrict__ z, double *__restrict__ scal) {
    __shared__ double red[32];
    if (scal[7] != 0.0) return;     // uniform over the grid: nobody reaches a grid barrier
    cg::grid_group grid = cg::this_grid();
    const double rho = scal[0];
    const int64_t n = npix * POL;
    const int64_t stride = (int64_t)gridDim.x * VB;
    // phase 1: p.q
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += stride) s = fma(p[i], q[i], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) g_part[2][blockIdx.x] = s;
    grid.sync();
    const double pq = sum_partials(g_part[2], gridDim.x, red);
    const double alpha = rho / pq;
    // phase 2: x, r, z and the two reductions
    double s0 = 0.0, s1 = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * VB + threadIdx.x; j < npix; j += stride) {
        double rv[POL], zv[POL];
#pragma unroll
        for (int k = 0; k < POL; ++k) {
            const int64_t i = POL * j + k;
            x[i] = fma(alpha, p[i], x[i]);
            rv[k] = fma(-alpha, q[i], r[i]);
            r[i] = rv[k];
        }
        bd_z<POL>(inv, j, rv, zv);
#pragma unroll
        for (int k = 0; k < POL; ++k) {
            z[POL * j + k] = zv[k];
            s0 = fma(rv[k], zv[k], s0);
            s1 = fma(rv[k], rv[k], s1);
        }
    }
    s0 = block_sum(s0, red);
    s1 = block_sum(s1, red);
    if (threadIdx.x == 0) { g_part[0][blockIdx.x] = s0; g_part[1][blockIdx.x] = s1; }
    grid.sync();
    const double rho_new = sum_partials(g_part[0], gridDim.x, red);
    const double rr = sum_partials(g_part[1], gridDim.x, red);
    const double beta = rho_new / rho;
    // phase 3: next search direction (same pixel -> thread mapping as phase 2: z is this thread's own)
    for (int64_t j = (int64_t)blockIdx.x * VB + threadIdx.x; j < npix; j += stride) {
#pragma unroll
        for (int k = 0; k < POL; ++k) {
            const int64_t i = POL * j + k;
            p[i] = fma(beta, p[i], z[i]);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[1] = rho;
        scal[0] = rho_new;
        scal[2] = pq;
        scal[3] = rr;
        scal[4] = alpha;
        scal[5] = beta;
        scal[8] += 1.0;
        scal[7] = (sqrt(rr) < scal[6]) ? 1.0 : 0.0;
    }
}

static int vgrid(int64_t n) {
    int64_t b = (n + VB - 1) / VB;
    int64_t cap = (int64_t)sm_count() * 4;
    if (cap > MAXP) cap = MAXP;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace cm2

using namespace cm2;

extern "C" int cm2_dot(const double *a, const double *b, int64_t n, double *out, cm2_stream_t stream) {
    CM2_REQUIRE(n >= 0, "n < 0");
    k_dot<<<vgrid(n), VB, 0, as_stream(stream)>>>(a, b, n, out);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_axpby(double alpha, const double *x, double beta, double *y, int64_t n, cm2_stream_t stream) {
    CM2_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return CM2_OK;
    k_axpby<<<vgrid(n), VB, 0, as_stream(stream)>>>(alpha, x, beta, y, n);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_pcg_reset(const double *r, int64_t n, double *scal, double atol, cm2_stream_t stream) {
    CM2_REQUIRE(n >= 0, "n < 0");
    k_reset<<<vgrid(n), VB, 0, as_stream(stream)>>>(r, n, scal, atol);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_pcg_update_p(const double *r, const double *z, double *p, int64_t n, double *scal,
                                cm2_stream_t stream) {
    CM2_REQUIRE(n >= 0, "n < 0");
    cudaStream_t st = as_stream(stream);
    k_rho<<<vgrid(n), VB, 0, st>>>(r, z, n, scal);
    CM2_LAUNCHED();
    k_update_p<<<vgrid(n), VB, 0, st>>>(z, p, n, scal);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_pcg_update_xr(const double *p, const double *q, double *x, double *r, int64_t n, double *scal,
                                 cm2_stream_t stream) {
    CM2_REQUIRE(n >= 0, "n < 0");
    cudaStream_t st = as_stream(stream);
    k_pq<<<vgrid(n), VB, 0, st>>>(p, q, n, scal);
    CM2_LAUNCHED();
    k_update_xr<<<vgrid(n), VB, 0, st>>>(p, q, x, r, n, scal);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_pcg_bd_reset(const double *inv, int64_t npix, int pol, double *r, double *z, double *scal,
                                double atol, double rtol, const double *b, double *x, double *p0, cm2_stream_t stream) {
    CM2_REQUIRE(npix >= 0 && pol >= 1 && pol <= 3, "bad npix/pol");
    CM2_REQUIRE(aligned(inv, 16), "inverse blocks must be 16-byte aligned");
    CM2_REQUIRE((b == nullptr) == (x == nullptr), "b and x go together");
    cudaStream_t st = as_stream(stream);
    const int g = vgrid(npix);
    if (pol == 1) k_bd_reset<1><<<g, VB, 0, st>>>(inv, npix, r, z, scal, atol, rtol, b, x, p0);
    else if (pol == 2) k_bd_reset<2><<<g, VB, 0, st>>>(inv, npix, r, z, scal, atol, rtol, b, x, p0);
    else k_bd_reset<3><<<g, VB, 0, st>>>(inv, npix, r, z, scal, atol, rtol, b, x, p0);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_pcg_bd_update_p(const double *z, double *p, int64_t n, double *scal, cm2_stream_t stream) {
    CM2_REQUIRE(n >= 0, "n < 0");
    k_update_p<<<vgrid(n), VB, 0, as_stream(stream)>>>(z, p, n, scal);
    CM2_LAUNCHED();
    return CM2_OK;
}

extern "C" int cm2_pcg_bd_update(const double *inv, int64_t npix, int pol, const double *p, const double *q, double *x,
                                 double *r, double *z, double *scal, cm2_stream_t stream) {
    CM2_REQUIRE(npix >= 0 && pol >= 1 && pol <= 3, "bad npix/pol");
    CM2_REQUIRE(aligned(inv, 16), "inverse blocks must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int64_t n = npix * pol;
    k_pq<<<vgrid(n), VB, 0, st>>>(p, q, n, scal);
    CM2_LAUNCHED();
    const int g = vgrid(npix);
    if (pol == 1) k_bd_update<1><<<g, VB, 0, st>>>(inv, npix, p, q, x, r, z, scal);
    else if (pol == 2) k_bd_update<2><<<g, VB, 0, st>>>(inv, npix, p, q, x, r, z, scal);
    else k_bd_update<3><<<g, VB, 0, st>>>(inv, npix, p, q, x, r, z, scal);
    CM2_LAUNCHED();
    return CM2_OK;
}

// test hook (cm2_pcg_bd_iter_refuse): make the cooperative launch fail the way a GPU that refuses
// cooperative launches does (too many CTAs for co-residency), to exercise the fallback
static int g_refuse_coop = 0;

template <int POL>
static int launch_bd_iter(const double *inv, int64_t npix, double *p, const double *q, double *x, double *r, double *z,
                          double *scal, cudaStream_t st) {
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bd_iter<POL>, VB, 0);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int64_t g = (int64_t)sm_count() * per_sm;            // all CTAs co-resident: grid barriers are safe
    if (g > MAXP) g = MAXP;
    const int64_t need = (npix + VB - 1) / VB;
    if (need < g) g = need < 1 ? 1 : need;
    if (g_refuse_coop) g = (int64_t)sm_count() * 64;     // more CTAs than can be co-resident: the launch is refused
    void *args[] = {(void *)&inv, (void *)&npix, (void *)&p, (void *)&q, (void *)&x, (void *)&r, (void *)&z, (void *)&scal};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_bd_iter<POL>, dim3((unsigned)g), dim3(VB), args, 0, st);
    if (e != cudaSuccess) {
        cudaGetLastError();      // clear the runtime's sticky last-error: the caller falls back to the 3-kernel tail
        return set_error(CM2_ERR_CUDA, "cm2_pcg_bd_iter: %s", cudaGetErrorString(e));
    }
    count_launch();
    return CM2_OK;
}

extern "C" int cm2_pcg_bd_iter_refuse(int on) {
    const int old = g_refuse_coop;
    g_refuse_coop = on ? 1 : 0;
    return old;
}

extern "C" int cm2_pcg_bd_iter(const double *inv, int64_t npix, int pol, double *p, const double *q, double *x,
                               double *r, double *z, double *scal, cm2_stream_t stream) {
    CM2_REQUIRE(npix >= 0 && pol >= 1 && pol <= 3, "bad npix/pol");
    CM2_REQUIRE(aligned(inv, 16), "inverse blocks must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    if (pol == 1) return launch_bd_iter<1>(inv, npix, p, q, x, r, z, scal, st);
    if (pol == 2) return launch_bd_iter<2>(inv, npix, p, q, x, r, z, scal, st);
    return launch_bd_iter<3>(inv, npix, p, q, x, r, z, scal, st);
}
