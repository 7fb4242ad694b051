// cosmomap2_b200 -- PCG vector work on the device (sm_100a).
//
// Restates the recurrence of scipy.sparse.linalg.cg (scipy/_isolve/iterative.py:405-431), which the
// reference calls at src/test_BD_precond_onto_real_data.py:47 and
// src/test_M2_precond_onto_real_data.py:117, with the scalars kept in an 8-double device
// workspace so an iteration needs no host round trip for alpha/beta:
//   scal[0]=rho  [1]=rho_prev  [2]=p.q  [3]=|r|^2  [4]=alpha  [5]=beta  [6],[7] caller-owned
//
// Reductions are deterministic: fixed per-thread order, fixed shuffle tree, per-CTA partials
// summed in CTA order by the last CTA to finish (ticket).  The partial/ticket scratch is a single
// device-global area: call these entry points from one stream at a time.
#include "cm2_common.cuh"

namespace cm2 {

constexpr int VB = 256;
constexpr int MAXP = 2048;  // max CTAs of a reduction kernel

__device__ double g_part[3][MAXP];
__device__ unsigned int g_ticket[3];

// CTA-level: write partial, last CTA reduces all partials in order; returns true in thread 0 of
// the last CTA with the total in *total
__device__ __forceinline__ bool finish_reduce(double v, int slot, double *red, double *total) {
    __shared__ bool last;
    const double t = block_sum(v, red);
    if (threadIdx.x == 0) {
        g_part[slot][blockIdx.x] = t;
        __threadfence();
        const unsigned int tk = atomicAdd(&g_ticket[slot], 1u);
        last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += VB) s += ((volatile double *)g_part[slot])[i];
    const double tot = block_sum(s, red);
    if (threadIdx.x == 0) {
        *total = tot;
        g_ticket[slot] = 0;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(VB) k_dot(const double *__restrict__ a, const double *__restrict__ b, int64_t n,
                                            double *__restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) s = fma(a[i], b[i], s);
    double tot;
    if (finish_reduce(s, 0, red, &tot)) *out = tot;
}

__global__ void __launch_bounds__(VB) k_axpby(double alpha, const double *__restrict__ x, double beta, double *__restrict__ y,
                                              int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        const double yi = beta == 0.0 ? 0.0 : beta * y[i];
        y[i] = fma(alpha, x[i], yi);
    }
}

// rho = r.z ; then (second kernel) p = z + (rho/rho_prev) p
__global__ void __launch_bounds__(VB) k_rho(const double *__restrict__ r, const double *__restrict__ z, int64_t n,
                                            double *__restrict__ scal, int first) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) s = fma(r[i], z[i], s);
    double tot;
    if (finish_reduce(s, 0, red, &tot)) {
        scal[0] = tot;
        scal[5] = first ? 0.0 : tot / scal[1];
    }
}

__global__ void __launch_bounds__(VB) k_update_p(const double *__restrict__ z, double *__restrict__ p, int64_t n,
                                                 const double *__restrict__ scal, int first) {
    const double beta = scal[5];
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB)
        p[i] = first ? z[i] : fma(beta, p[i], z[i]);
}

__global__ void __launch_bounds__(VB) k_pq(const double *__restrict__ p, const double *__restrict__ q, int64_t n,
                                           double *__restrict__ scal) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) s = fma(p[i], q[i], s);
    double tot;
    if (finish_reduce(s, 1, red, &tot)) {
        scal[2] = tot;
        scal[4] = scal[0] / tot;
    }
}

__global__ void __launch_bounds__(VB) k_update_xr(const double *__restrict__ p, const double *__restrict__ q,
                                                  double *__restrict__ x, double *__restrict__ r, int64_t n,
                                                  double *__restrict__ scal) {
    __shared__ double red[32];
    const double alpha = scal[4];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VB + threadIdx.x; i < n; i += (int64_t)gridDim.x * VB) {
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, q[i], r[i]);
        r[i] = ri;
        s = fma(ri, ri, s);
    }
    double tot;
    if (finish_reduce(s, 2, red, &tot)) {
        scal[3] = tot;
        scal[1] = scal[0];
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

extern "C" int cm2_pcg_update_p(const double *r, const double *z, double *p, int64_t n, double *scal, int first,
                                cm2_stream_t stream) {
    CM2_REQUIRE(n >= 0, "n < 0");
    cudaStream_t st = as_stream(stream);
    k_rho<<<vgrid(n), VB, 0, st>>>(r, z, n, scal, first);
    CM2_LAUNCHED();
    k_update_p<<<vgrid(n), VB, 0, st>>>(z, p, n, scal, first);
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
