// Width-generic pipeline (see alg_generic.cuh): host orchestration + the non-tensor-product kernels.
// Chain rule as stated (and checked against autograd) in oracle/analytic_numpy.py.
#include <algorithm>
#include <cstdlib>
#include <string>

#include "alg_generic.cuh"

namespace alg {

GenTpDims gen_tp_dims(int L, char kind) {
  return L == 1 ? gen_tp_dims_L1(kind) : (L == 2 ? gen_tp_dims_L2(kind) : gen_tp_dims_L3(kind));
}
cudaError_t gen_tp_launch(int L, char kind, bool first, bool backward, const GenTp& a, cudaStream_t st) {
  return L == 1 ? gen_tp_launch_L1(kind, first, backward, a, st)
                : (L == 2 ? gen_tp_launch_L2(kind, first, backward, a, st) : gen_tp_launch_L3(kind, first, backward, a, st));
}

namespace {

struct EdgeGeo { float x, y, z, r, rc, u, dudr; int zi, zj; };
__device__ __forceinline__ EdgeGeo gen_geo(const float4 rv, const GenTables& tb, float p) {
  EdgeGeo g;
  const int zz = __float_as_int(rv.w);
  g.zi = zz & 255; g.zj = zz >> 8;
  g.r = sqrtf(rv.x * rv.x + rv.y * rv.y + rv.z * rv.z);
  g.x = rv.x / g.r; g.y = rv.y / g.r; g.z = rv.z / g.r;
  g.rc = tb.rc[g.zi * MAXT + g.zj];
  float dudx;
  poly_cutoff(g.r / g.rc, p, g.u, dudx);
  g.dudr = dudx / g.rc;
  return g;
}

// per edge: Y, u and the two-body MLP input (one-hot Z_i | one-hot Z_j | bessel * u)
template <int L>
__global__ void k_gen_geom(int n, int e0, const float4* __restrict__ rvec, const GenTables tb, float p, int T, int B,
                           float* __restrict__ Y, float* __restrict__ u, float* __restrict__ IN0) {
  constexpr int NSH = (L + 1) * (L + 1);
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const EdgeGeo g = gen_geo(rvec[e0 + q], tb, p);
  float Yl[NSH];
  sph_harm<L>(g.x, g.y, g.z, Yl);
#pragma unroll
  for (int lm = 0; lm < NSH; ++lm) Y[(size_t)q * NSH + lm] = Yl[lm];
  u[q] = g.u;
  const int K0 = 2 * T + B;
  float* in = IN0 + (size_t)q * K0;
  for (int t = 0; t < T; ++t) { in[t] = t == g.zi ? 1.f : 0.f; in[T + t] = t == g.zj ? 1.f : 0.f; }
  const float pref = sqrtf(2.0f / g.rc);
  const float xr = g.r / g.rc;
  for (int b = 0; b < B; ++b) in[2 * T + b] = pref * sinf((float)(b + 1) * (3.14159265358979323846f * xr)) / g.r * g.u;
}

// GEMM epilogues: what happens to an accumulator v of element (row gm, column gj) of C[n][N]
enum { EPI_STORE = 0, EPI_ADD = 1, EPI_SILU = 2, EPI_MUL = 3 };
// EPI_SILU: C = c*silu(v), D[gm][gj] = its derivative (D dense, leading dimension N);  EPI_MUL: C = v * D[gm][gj]
__device__ __forceinline__ void gemm_epilogue(int mode, float* __restrict__ C, int ldc, float* __restrict__ D, int N, int gm, int gj, float v) {
  float* dst = C + (size_t)gm * ldc + gj;
  if (mode == EPI_STORE) *dst = v;
  else if (mode == EPI_ADD) *dst += v;
  else if (mode == EPI_SILU) { float dd; *dst = silu_act(v, dd); D[(size_t)gm * N + gj] = dd; }
  else *dst = v * D[(size_t)gm * N + gj];
}

// C[n][N] (+)= A[n][K] . W[K][N]   (row-major, leading dimensions lda / ldw / ldc); k ascending -> deterministic
constexpr int GB = 64, GK = 16;
__global__ void __launch_bounds__(256) k_gen_gemm(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                  float* __restrict__ C, int ldc, int n, int K, int N, int mode, float* __restrict__ D) {
  __shared__ float As[GK][GB + 1];
  __shared__ float Ws[GK][GB];
  const int m0 = blockIdx.x * GB, j0 = blockIdx.y * GB;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  float c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = t + 256 * i;
      const int r = idx >> 4, kk = idx & 15;
      const int gm = m0 + r, gk = k0 + kk;
      As[kk][r] = (gm < n && gk < K) ? A[(size_t)gm * lda + gk] : 0.f;
      const int k2 = idx >> 6, j = idx & 63;
      const int gk2 = k0 + k2, gj = j0 + j;
      Ws[k2][j] = (gk2 < K && gj < N) ? W[(size_t)gk2 * ldw + gj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a4[4], w4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a4[i] = As[kk][ty * 4 + i]; w4[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = fmaf(a4[i], w4[j], c[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = j0 + tx * 4 + j;
      if (gj >= N) continue;
      gemm_epilogue(mode, C, ldc, D, N, gm, gj, c[i][j]);
    }
  }
}

// The same GEMM on the tensor cores: warp-level mma.sync m16n8k8 TF32 with the 3xTF32 split (hi*hi + hi*lo + lo*hi, small
// terms first) for fp32-level accuracy -- the split the tcgen05 pipeline uses (umma.cuh), here with operands of any shape
// staged through shared memory.  Block 128 x 64, 8 warps of 32 x 32.  Fixed summation order -> deterministic.
constexpr int MB = 128, NB = 64, KB = 16;
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void __launch_bounds__(256) k_gen_gemm_mma(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                                      float* __restrict__ C, int ldc, int n, int K, int N, int mode, float* __restrict__ D) {
  __shared__ float As[MB][KB + 4];
  __shared__ float Bs[KB][NB + 8];
  const int m0 = blockIdx.x * MB, j0 = blockIdx.y * NB;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, g = lane >> 2, tig = lane & 3;
  const int wm = warp >> 1, wn = warp & 1;
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
  float pa[8], pb[4];                                 // next K-slab, fetched while the current one is multiplied
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = t + 256 * i;
      const int gm = m0 + (idx >> 4), gk = k0 + (idx & 15);
      pa[i] = (gm < n && gk < K) ? A[(size_t)gm * lda + gk] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = t + 256 * i;
      const int gk = k0 + (idx >> 6), gj = j0 + (idx & 63);
      pb[i] = (gk < K && gj < N) ? W[(size_t)gk * ldw + gj] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += KB) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int idx = t + 256 * i; As[idx >> 4][idx & 15] = pa[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int idx = t + 256 * i; Bs[idx >> 6][idx & 63] = pb[i]; }
    __syncthreads();
    if (k0 + KB < K) fetch(k0 + KB);
#pragma unroll
    for (int ks = 0; ks < KB; ks += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = wm * 32 + mt * 16 + g;
        split_tf32(As[r][ks + tig], ah[mt][0], al[mt][0]);
        split_tf32(As[r + 8][ks + tig], ah[mt][1], al[mt][1]);
        split_tf32(As[r][ks + tig + 4], ah[mt][2], al[mt][2]);
        split_tf32(As[r + 8][ks + tig + 4], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = wn * 32 + nt * 8 + g;
        split_tf32(Bs[ks + tig][c], bh[nt][0], bl[nt][0]);
        split_tf32(Bs[ks + tig + 4][c], bh[nt][1], bl[nt][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma_tf32(acc[mt][nt], al[mt], bh[nt]);
          mma_tf32(acc[mt][nt], ah[mt], bl[nt]);
          mma_tf32(acc[mt][nt], ah[mt], bh[nt]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int gm = m0 + wm * 32 + mt * 16 + g + (q >> 1) * 8;
        const int gj = j0 + wn * 32 + nt * 8 + 2 * tig + (q & 1);
        if (gm < n && gj < N) gemm_epilogue(mode, C, ldc, D, N, gm, gj, acc[mt][nt][q]);
      }
}

__global__ void k_gen_zero(long total, float* __restrict__ x) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) x[i] = 0.f;
}
// out[q][s] = a * base[q][s] + b * m[q][s] * u[q]   (base may be null)
__global__ void k_gen_mix(int n, int S, const float* __restrict__ base, const float* __restrict__ m, const float* __restrict__ u,
                          float a, float b, float* __restrict__ out, float* __restrict__ out2, int ld2) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)n * S) return;
  const int q = (int)(i / S);
  const float v = (base ? a * base[i] : 0.f) + b * m[i] * u[q];
  out[i] = v;
  if (out2) out2[(size_t)q * ld2 + (i - (long)q * S)] = v;
}
// backward of out = a*x + b*m*u:  dm = b*dout*u, du += sum_s b*dout*m, dx(in place) = a*dout.   One warp per edge.
__global__ void k_gen_mix_bwd(int n, int S, float* __restrict__ dX, const float* __restrict__ m, const float* __restrict__ u,
                              float a, float b, float* __restrict__ dM, float* __restrict__ du) {
  const int q = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (q >= n) return;
  const float uq = u[q];
  float acc = 0.f;
  for (int s = lane; s < S; s += 32) {
    const size_t i = (size_t)q * S + s;
    const float dxt = b * dX[i];
    acc += dxt * m[i];
    dM[i] = dxt * uq;
    dX[i] = a * dX[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) du[q] += acc;
}
// Gamma_c[lm*U+u] = inv * sum_{e in N(c)} w[e][l*U+u] * Y[e][lm]   (edges in list order)
__global__ void k_gen_gamma(int c0, int e0, const int* __restrict__ rowptr, int U, int NSH, int ENVW, const float* __restrict__ w,
                            const float* __restrict__ Y, float inv, float* __restrict__ gamma) {
  const int c = c0 + blockIdx.x;
  const int q0 = rowptr[c] - e0, q1 = rowptr[c + 1] - e0;
  const int F = NSH * U;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const int lm = f / U, uu = f % U;
    const int col = lsel(lm) * U + uu;
    float acc = 0.f;
    for (int q = q0; q < q1; ++q) acc += w[(size_t)q * ENVW + col] * Y[(size_t)q * NSH + lm];
    gamma[(size_t)blockIdx.x * F + f] = acc * inv;
  }
}
// dGamma_c[lm*U+u] = sum_{e in N(c)} dG_e[lm][e][u]
__global__ void k_gen_dgamma(int c0, int e0, const int* __restrict__ rowptr, int U, int NSH, size_t NU, const float* __restrict__ dge,
                             float* __restrict__ dgamma) {
  const int c = c0 + blockIdx.x;
  const int q0 = rowptr[c] - e0, q1 = rowptr[c + 1] - e0;
  const int F = NSH * U;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const int lm = f / U, uu = f % U;
    float acc = 0.f;
    for (int q = q0; q < q1; ++q) acc += dge[(size_t)lm * NU + (size_t)q * U + uu];
    dgamma[(size_t)blockIdx.x * F + f] = acc;
  }
}
// Backward of an outer product with the spherical harmonics, one warp per edge (lanes over channels, everything coalesced):
//   dw[e][l*U+u] = inv * sum_{m in l} dT[lm][u] * Y[e][lm]        dY[e][lm] += inv * sum_u dT[lm][u] * w[e][l*U+u]
// MODE 0: Gamma = inv * sum_e w (x) Y  -> dT = dGamma of the edge's centre ([lm*U+u]);
// MODE 1: V^0 = w0 (x) Y               -> dT = dV^0 of the edge itself (component-major [lm][e][u]), inv = 1.
template <int L, int MODE>
__global__ void k_gen_outer_bwd(int n, int e0, int c0, const int* __restrict__ edge_c, int U, const float* __restrict__ dsrc,
                                const float* __restrict__ Y, const float* __restrict__ w, float inv, float* __restrict__ dw,
                                float* __restrict__ dY) {
  constexpr int NSH = (L + 1) * (L + 1);
  const int q = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (q >= n) return;
  const int ENVW = (L + 1) * U;
  const size_t NU = (size_t)n * U;
  const float* dT = MODE == 0 ? dsrc + (size_t)(edge_c[e0 + q] - c0) * NSH * U : dsrc + (size_t)q * U;
  const size_t cs = MODE == 0 ? (size_t)U : NU;               // stride between components lm
  float y[NSH], dyp[NSH];
#pragma unroll
  for (int lm = 0; lm < NSH; ++lm) { y[lm] = Y[(size_t)q * NSH + lm]; dyp[lm] = 0.f; }
  for (int u = lane; u < U; u += 32) {
#pragma unroll
    for (int l = 0; l <= L; ++l) {
      const float wv = w[(size_t)q * ENVW + l * U + u];
      float acc = 0.f;
#pragma unroll
      for (int lm = l * l; lm < (l + 1) * (l + 1); ++lm) {
        const float d = dT[(size_t)lm * cs + u];
        acc += d * y[lm];
        dyp[lm] += d * wv;
      }
      dw[(size_t)q * ENVW + l * U + u] = acc * inv;
    }
  }
#pragma unroll
  for (int lm = 0; lm < NSH; ++lm) {
    float v = dyp[lm];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) dY[(size_t)q * NSH + lm] += v * inv;
  }
}
// readout second layer: E_e = ro1 . a ;  dz (in place over act') = gscale[Z_i] * ro1 * act'
__global__ void k_gen_readout(int n, int e0, int R, const float* __restrict__ ar, float* __restrict__ dr, const float* __restrict__ ro1,
                              const float4* __restrict__ rvec, const GenTables tb, float* __restrict__ Ee, float* __restrict__ edge_energy) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int zi = __float_as_int(rvec[e0 + q].w) & 255;
  const float ge = tb.gscale[zi];
  float ee = 0.f;
  for (int r = 0; r < R; ++r) {
    const float wq = __ldg(ro1 + r);
    ee += wq * ar[(size_t)q * R + r];
    dr[(size_t)q * R + r] *= ge * wq;
  }
  Ee[q] = ee;
  if (edge_energy) edge_energy[e0 + q] = ee;
}
// raw per-centre energy sums (double, edges in order)
__global__ void k_gen_esum(int c0, int nc, int e0, const int* __restrict__ rowptr, const float* __restrict__ Ee, double* __restrict__ esum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  const int c = c0 + i;
  double acc = 0.0;
  for (int q = rowptr[c] - e0; q < rowptr[c + 1] - e0; ++q) acc += (double)Ee[q];
  esum[c] = acc;
}
// geometry backward + force / virial accumulation (same conventions as k_b0, allegro_kernels.cuh)
template <int L>
__global__ void k_gen_force(int n, int e0, const float4* __restrict__ rvec, const GenTables tb, float p, int T, int B,
                            const float* __restrict__ du, const float* __restrict__ dIN0, const float* __restrict__ dY,
                            const int* __restrict__ edge_j, const int* __restrict__ edge_c, const int* __restrict__ ilist,
                            unsigned long long* __restrict__ facc, unsigned long long* __restrict__ vacc, float* __restrict__ edge_grad) {
  constexpr int NSH = (L + 1) * (L + 1);
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  float vir[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (q < n) {
    const int e = e0 + q;
    const EdgeGeo g = gen_geo(rvec[e], tb, p);
    const float pref = sqrtf(2.0f / g.rc);
    const float xr = g.r / g.rc;
    float dr = du[q] * g.dudr;
    const float* dbu = dIN0 + (size_t)q * (2 * T + B) + 2 * T;
    for (int b = 0; b < B; ++b) {
      const float kn = (float)(b + 1) * 3.14159265358979323846f;
      float sn, cs;
      sincosf(kn * xr, &sn, &cs);
      const float bes = pref * sn / g.r;
      const float dbes = pref * (kn / g.rc * cs / g.r - sn / (g.r * g.r));
      dr += dbu[b] * (dbes * g.u + bes * g.dudr);
    }
    float dYt[NSH];
#pragma unroll
    for (int lm = 0; lm < NSH; ++lm) dYt[lm] = dY[(size_t)q * NSH + lm];
    float qx, qy, qz;
    sph_harm_vjp<L>(g.x, g.y, g.z, dYt, qx, qy, qz);
    const float nq = g.x * qx + g.y * qy + g.z * qz;
    const float ir = 1.0f / g.r;
    const float gx = dr * g.x + (qx - g.x * nq) * ir;
    const float gy = dr * g.y + (qy - g.y * nq) * ir;
    const float gz = dr * g.z + (qz - g.z * nq) * ir;
    if (edge_grad) { edge_grad[3 * (size_t)e + 0] = gx; edge_grad[3 * (size_t)e + 1] = gy; edge_grad[3 * (size_t)e + 2] = gz; }
    const int j = edge_j[e], i = ilist[edge_c[e]];
    const float gv[3] = {gx, gy, gz};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atomicAdd(facc + 3 * (size_t)j + k, (unsigned long long)__double2ll_rn(-(double)gv[k] * FIX_SCALE));
      atomicAdd(facc + 3 * (size_t)i + k, (unsigned long long)__double2ll_rn((double)gv[k] * FIX_SCALE));
    }
    const float rx = g.x * g.r, ry = g.y * g.r, rz = g.z * g.r;
    vir[0] = -rx * gx; vir[1] = -ry * gy; vir[2] = -rz * gz;
    vir[3] = -0.5f * (rx * gy + ry * gx); vir[4] = -0.5f * (rx * gz + rz * gx); vir[5] = -0.5f * (ry * gz + rz * gy);
  }
  if (vacc) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      long long v = __double2ll_rn((double)vir[k] * VIR_SCALE);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v != 0) atomicAdd(vacc + k, (unsigned long long)v);
    }
  }
}

// ---- workspace plan -------------------------------------------------------------------------------------------------
struct Plan {
  long n, nc;
  int NSH, ENVW, SIN, K0, F, maxw, maxd;
  int vdim[3];                       // components per channel of V^k (k >= 1)
  size_t Y, u, du, dY, IN0, D[4][GEN_MAXLIN], M[4], X[4], W0e, Wk[3], V[3], gamma[3], dgamma, IN, SC[3], Dr, Ee, dX, dV[2], dGe, total;
};
Plan make_plan(const GenModel& m, long n, long nc) {
  Plan p{};
  p.n = n; p.nc = nc;
  p.NSH = (m.L + 1) * (m.L + 1); p.ENVW = (m.L + 1) * m.U; p.K0 = 2 * m.T + m.B; p.F = p.NSH * m.U;
  p.SIN = m.S + gen_tp_dims(m.L, 'A').n0 * m.U;
  p.maxw = std::max({p.SIN, m.H, m.S, p.ENVW, p.K0, m.R});
  p.maxd = p.NSH;
  for (int k = 1; k < m.nl; ++k) { p.vdim[k] = gen_tp_dims(m.L, m.layer[k].kind).din; p.maxd = std::max(p.maxd, p.vdim[k]); }
  size_t off = 0;
  auto take = [&](size_t cnt) { const size_t o = off; off += (cnt + 63) / 64 * 64; return o; };
  const size_t N = (size_t)n;
  p.Y = take(N * p.NSH); p.u = take(N); p.du = take(N); p.dY = take(N * p.NSH); p.IN0 = take(N * p.K0);
  for (int st = 0; st <= m.nl; ++st) {
    for (int i = 0; i < m.depth; ++i) p.D[st][i] = take(N * m.H);
    p.M[st] = take(N * m.S);
    p.X[st] = take(N * m.S);
  }
  p.W0e = take(N * p.ENVW);
  for (int k = 0; k < m.nl; ++k) { p.Wk[k] = take(N * p.ENVW); p.gamma[k] = take((size_t)nc * p.F); }
  for (int k = 1; k < m.nl; ++k) p.V[k] = take(N * m.U * p.vdim[k]);
  p.dgamma = take((size_t)nc * p.F);
  p.IN = take(N * std::max(p.SIN, p.K0));
  for (int i = 0; i < 3; ++i) p.SC[i] = take(N * p.maxw);
  p.Dr = take(N * m.R); p.Ee = take(N); p.dX = take(N * m.S);
  for (int i = 0; i < 2; ++i) p.dV[i] = take(N * m.U * p.maxd);
  p.dGe = take(N * m.U * p.NSH);
  p.total = off;
  return p;
}

struct Runner {
  const GenModel& m; float* ws; Plan p; cudaStream_t st; long launches = 0; cudaError_t err = cudaSuccess;
  int n;
  unsigned blocks(long total, int tpb = 256) const { return (unsigned)((total + tpb - 1) / tpb); }
  void check() { if (err == cudaSuccess) err = cudaGetLastError(); ++launches; }
  void gemm(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int K, int N, int mode, float* D = nullptr) {
    if (n == 0 || N <= 0) return;
    static const bool ffma = [] { const char* e = getenv("ALG_GENERIC_GEMM"); return e && std::string(e) == "ffma"; }();   // debugging aid
    if (ffma) k_gen_gemm<<<dim3((n + GB - 1) / GB, (N + GB - 1) / GB), 256, 0, st>>>(A, lda, W, ldw, C, ldc, n, K, N, mode, D);
    else k_gen_gemm_mma<<<dim3((n + MB - 1) / MB, (N + NB - 1) / NB), 256, 0, st>>>(A, lda, W, ldw, C, ldc, n, K, N, mode, D);
    check();
  }
  // keeps act' of hidden layer i of stage `stage` in D[stage][i]; result (last linear, no activation) -> out [n][dims[nlin]]
  void mlp_fwd(const GenMLP& mlp, const float* in, int ldin, int stage, float* out) {
    const float* cur = in; int ld = ldin;
    for (int i = 0; i < mlp.nlin; ++i) {
      const bool last = i == mlp.nlin - 1;
      float* z = last ? out : ws + p.SC[i & 1];
      gemm(cur, ld, mlp.w[i], mlp.dims[i + 1], z, mlp.dims[i + 1], mlp.dims[i], mlp.dims[i + 1], last ? EPI_STORE : EPI_SILU,
           last ? nullptr : ws + p.D[stage][i]);
      cur = z; ld = mlp.dims[i + 1];
    }
  }
  // dout [n][dims[nlin]] (in SC[2]) -> din [n][dims[0]].  With dx != nullptr the first `split` input columns are ADDED to
  // dx [n][split] instead (the latent part of a layer's MLP input), the rest goes to din at the same column offset.
  void mlp_bwd(const GenMLP& mlp, const float* dout, int stage, float* din, float* dx = nullptr, int split = 0) {
    const float* cur = dout;
    int flip = 0;
    for (int i = mlp.nlin - 1; i >= 0; --i) {
      float* o = i == 0 ? din : ws + p.SC[flip];
      flip ^= 1;
      const int K = mlp.dims[i + 1], N = mlp.dims[i];
      if (i == 0 && dx) {
        gemm(cur, K, mlp.wt[0], N, dx, split, K, split, EPI_ADD);
        gemm(cur, K, mlp.wt[0] + split, N, o + split, N, K, N - split, EPI_STORE);
      } else {
        gemm(cur, K, mlp.wt[i], N, o, N, K, N, i > 0 ? EPI_MUL : EPI_STORE, i > 0 ? ws + p.D[stage][i - 1] : nullptr);
      }
      cur = o;
    }
  }
};

template <int L> void launch_geom(Runner& r, const GenTables& tb, const GenEdges& g, int e0) {
  k_gen_geom<L><<<r.blocks(r.n, 128), 128, 0, r.st>>>(r.n, e0, g.rvec, tb, r.m.p, r.m.T, r.m.B, r.ws + r.p.Y, r.ws + r.p.u, r.ws + r.p.IN0);
  r.check();
}
template <int L> void launch_force(Runner& r, const GenTables& tb, const GenEdges& g, int e0, const float* dIN0) {
  k_gen_force<L><<<r.blocks(r.n, 128), 128, 0, r.st>>>(r.n, e0, g.rvec, tb, r.m.p, r.m.T, r.m.B, r.ws + r.p.du, dIN0, r.ws + r.p.dY,
                                                         g.edge_j, g.edge_c, g.ilist, g.facc, g.vacc, g.edge_grad);
  r.check();
}

void launch_outer_bwd(Runner& r, int L, int mode, int e0, int c0, const int* edge_c, const float* dsrc, const float* w, float inv, float* dw) {
  const unsigned b = r.blocks((long)r.n * 32, 256);
  const float* Y = r.ws + r.p.Y;
  float* dY = r.ws + r.p.dY;
  const int U = r.m.U;
#define ALG_OB(LV, MV) k_gen_outer_bwd<LV, MV><<<b, 256, 0, r.st>>>(r.n, e0, c0, edge_c, U, dsrc, Y, w, inv, dw, dY)
  if (mode == 0) { if (L == 1) ALG_OB(1, 0); else if (L == 2) ALG_OB(2, 0); else ALG_OB(3, 0); }
  else { if (L == 1) ALG_OB(1, 1); else if (L == 2) ALG_OB(2, 1); else ALG_OB(3, 1); }
#undef ALG_OB
  r.check();
}

}  // namespace

size_t gen_work_floats(const GenModel& m, long n, long nc) { return make_plan(m, n, nc).total; }

cudaError_t gen_run_chunk(const GenModel& m, const GenTables& tb, const GenEdges& g, int c0, int c1, int e0, int e1, float* ws,
                          cudaStream_t st, long* launches) {
  const int n = e1 - e0, nc = c1 - c0;
  if (n <= 0 || nc <= 0) return cudaSuccess;
  Runner r{m, ws, make_plan(m, n, nc), st};
  r.n = n;
  const Plan& p = r.p;
  const int S = m.S, U = m.U, L = m.L, nl = m.nl;
  auto W = [&](size_t off) { return ws + off; };
  // ---- forward
  if (L == 1) launch_geom<1>(r, tb, g, e0); else if (L == 2) launch_geom<2>(r, tb, g, e0); else launch_geom<3>(r, tb, g, e0);
  k_gen_zero<<<r.blocks(n), 256, 0, st>>>(n, W(p.du)); r.check();
  k_gen_zero<<<r.blocks((long)n * p.NSH), 256, 0, st>>>((long)n * p.NSH, W(p.dY)); r.check();
  r.mlp_fwd(m.two, W(p.IN0), p.K0, 0, W(p.M[0]));
  k_gen_mix<<<r.blocks((long)n * S), 256, 0, st>>>(n, S, nullptr, W(p.M[0]), W(p.u), 0.f, 1.f, W(p.X[0]), W(p.IN), p.SIN); r.check();
  r.gemm(W(p.X[0]), S, m.emb, p.ENVW, W(p.W0e), p.ENVW, S, p.ENVW, EPI_STORE);
  GenTp tp{};
  tp.n = n; tp.U = U; tp.S = S; tp.e0 = e0; tp.c0 = c0; tp.ldin = p.SIN; tp.envw = p.ENVW; tp.edge_c = g.edge_c; tp.Y = W(p.Y);
  for (int k = 0; k < nl; ++k) {
    const GenLayer& lw = m.layer[k];
    r.gemm(W(p.X[k]), S, lw.env, p.ENVW, W(p.Wk[k]), p.ENVW, S, p.ENVW, EPI_STORE);
    k_gen_gamma<<<nc, 256, 0, st>>>(c0, e0, g.rowptr, U, p.NSH, p.ENVW, W(p.Wk[k]), W(p.Y), m.inv_sqrt_n, W(p.gamma[k])); r.check();
    GenTp a = tp;
    a.vin = k == 0 ? W(p.W0e) : W(p.V[k]); a.gamma = W(p.gamma[k]); a.omega_t = lw.omega_t;
    a.vout = k < nl - 1 ? W(p.V[k + 1]) : nullptr; a.IN = W(p.IN);
    if (r.err == cudaSuccess) r.err = gen_tp_launch(L, lw.kind, k == 0, false, a, st);
    ++r.launches;
    r.mlp_fwd(lw.mlp, W(p.IN), p.SIN, k + 1, W(p.M[k + 1]));
    k_gen_mix<<<r.blocks((long)n * S), 256, 0, st>>>(n, S, W(p.X[k]), W(p.M[k + 1]), W(p.u), lw.a, lw.b, W(p.X[k + 1]),
                                                     k < nl - 1 ? W(p.IN) : nullptr, p.SIN); r.check();   // x^{k+1} is also the head of the next MLP input
  }
  // ---- readout, energies
  float* ar = W(p.SC[0]);
  r.gemm(W(p.X[nl]), S, m.ro0, m.R, ar, m.R, S, m.R, EPI_SILU, W(p.Dr));
  k_gen_readout<<<r.blocks(n, 128), 128, 0, st>>>(n, e0, m.R, ar, W(p.Dr), m.ro1, g.rvec, tb, W(p.Ee), g.edge_energy); r.check();
  k_gen_esum<<<r.blocks(nc, 128), 128, 0, st>>>(c0, nc, e0, g.rowptr, W(p.Ee), g.esum); r.check();
  // ---- backward
  r.gemm(W(p.Dr), m.R, m.ro0_t, S, W(p.dX), S, m.R, S, EPI_STORE);
  int cur = 0;                                       // dV[cur] = dE/dV^{k+1} (valid for k < nl-1)
  for (int k = nl - 1; k >= 0; --k) {
    const GenLayer& lw = m.layer[k];
    float* dM = W(p.SC[2]);
    k_gen_mix_bwd<<<r.blocks((long)n * 32), 256, 0, st>>>(n, S, W(p.dX), W(p.M[k + 1]), W(p.u), lw.a, lw.b, dM, W(p.du)); r.check();
    r.mlp_bwd(lw.mlp, dM, k + 1, W(p.IN), W(p.dX), S);   // dIN = (dx | ds): dx added to dX, ds -> columns S.. of IN
    GenTp a = tp;
    a.vin = k == 0 ? W(p.W0e) : W(p.V[k]); a.gamma = W(p.gamma[k]); a.omega_t = lw.omega_t; a.IN = W(p.IN);
    a.dvout = k < nl - 1 ? W(p.dV[cur]) : nullptr; a.dvin = W(p.dV[cur ^ 1]); a.dge = W(p.dGe);
    if (r.err == cudaSuccess) r.err = gen_tp_launch(L, lw.kind, k == 0, true, a, st);
    ++r.launches;
    cur ^= 1;                                        // dV[cur] = dE/dV^k
    k_gen_dgamma<<<nc, 256, 0, st>>>(c0, e0, g.rowptr, U, p.NSH, (size_t)n * U, W(p.dGe), W(p.dgamma)); r.check();
    float* dw = W(p.SC[2]);
    launch_outer_bwd(r, L, 0, e0, c0, g.edge_c, W(p.dgamma), W(p.Wk[k]), m.inv_sqrt_n, dw);
    r.gemm(dw, p.ENVW, lw.env_t, S, W(p.dX), S, p.ENVW, S, EPI_ADD);
    if (k == 0) {
      launch_outer_bwd(r, L, 1, e0, c0, g.edge_c, W(p.dV[cur]), W(p.W0e), 1.f, dw);
      r.gemm(dw, p.ENVW, m.emb_t, S, W(p.dX), S, p.ENVW, S, EPI_ADD);
    }
  }
  // two-body: x0 = m0 * u
  float* dM0 = W(p.SC[2]);
  k_gen_mix_bwd<<<r.blocks((long)n * 32), 256, 0, st>>>(n, S, W(p.dX), W(p.M[0]), W(p.u), 0.f, 1.f, dM0, W(p.du)); r.check();
  float* dIN0 = W(p.IN);
  r.mlp_bwd(m.two, dM0, 0, dIN0);
  if (L == 1) launch_force<1>(r, tb, g, e0, dIN0); else if (L == 2) launch_force<2>(r, tb, g, e0, dIN0); else launch_force<3>(r, tb, g, e0, dIN0);
  if (launches) *launches += r.launches;
  return r.err;
}

}  // namespace alg
