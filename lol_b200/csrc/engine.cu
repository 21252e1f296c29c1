// engine.cu -- the generic pass engine: any index m, any ring, one CTA per
// (ring element, RNS limb) with the element resident in shared memory (or in a
// per-CTA global workspace when it does not fit), one HBM read and one HBM write
// per transform.  The fused kernels (fused_*.cu) overtake it for the benchmark
// shapes; this is the path that makes every (m, q) of the reference work.
//
// Pass semantics follow lol-cpp one to one (file:line in lolb_internal.cuh's
// PassKind); the thread mapping does not: dense passes are one thread per
// OUTPUT coefficient, out of place between two buffers; diagonal passes one
// thread per coefficient; line passes (L, G, ...) one thread per line.
#include "lolb_internal.cuh"
#include "rings.cuh"

namespace lolb {

// ------------------------------------------------------------------ pass bodies

// I_L (x) DFT_2 (x) I_R in place: thread per butterfly (crt.cpp:137-149)
template <class R>
__device__ __forceinline__ void pass_dft2(const R& ring, const Pass& ps, typename R::T* cur, int n)
{
  const int Rr = ps.R;
  for (int b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
    int r = b % Rr, blk = b / Rr;
    int i0 = blk * 2 * Rr + r;
    typename R::T u = cur[i0], t = cur[i0 + Rr];
    cur[i0] = ring.add(u, t);
    cur[i0 + Rr] = ring.sub(u, t);
  }
}

// dense passes, thread per output, cur -> alt.  Root index walks (row*col) mod p incrementally.
template <class R, class W>
__device__ __forceinline__ void pass_dense(const R& ring, const Pass& ps, const typename R::T* cur,
                                           typename R::T* alt, int n, const W* tab)
{
  typedef typename R::T T;
  const int p = ps.p, d = ps.d, Rr = ps.R, rs = ps.rustride;
  const W* ru = tab + ps.tab;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int r = j % Rr, t = j / Rr;
    int row = t % d, blk = t / d;
    const T* in = cur + (size_t)blk * d * Rr + r;
    T acc = ring.zero();
    if (ps.kind == PASS_DFT) {                       // sum_col in[col] * ru[(row*col)%p]
      int idx = 0;
      for (int col = 0; col < p; col++) {
        acc = ring.add(acc, ring.mul(in[(size_t)col * Rr], ru[(size_t)idx * rs]));
        idx += row; if (idx >= p) idx -= p;
      }
    } else if (ps.kind == PASS_CRT) {                // sum_col in[col] * ru[((row+1)*col)%p]
      int idx = 0, step = row + 1;
      for (int col = 0; col < p - 1; col++) {
        acc = ring.add(acc, ring.mul(in[(size_t)col * Rr], ru[(size_t)idx * rs]));
        idx += step; if (idx >= p) idx -= p;
      }
    } else {                                         // PASS_CRTINV: sum_col in[col]*(ru[(row*(col+1))%p] ) - shift
      int idx = row;
      T shift = ring.zero();
      for (int col = 0; col < p - 1; col++) {
        T v = in[(size_t)col * Rr];
        acc = ring.add(acc, ring.mul(v, ru[(size_t)idx * rs]));
        shift = ring.add(shift, ring.mul(v, ru[(size_t)(p - col - 1) * rs]));
        idx += row; if (idx >= p) idx -= p;
      }
      acc = ring.sub(acc, shift);
    }
    alt[j] = acc;
  }
}

template <class R, class W>
__device__ __forceinline__ void pass_diag(const R& ring, const Pass& ps, typename R::T* cur, int n, const W* tab)
{
  const int d = ps.d, Rr = ps.R;
  const W* tw = tab + ps.tab;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int pos = (j / Rr) % d;
    cur[j] = ring.mul(cur[j], tw[pos]);
  }
}

// line passes: thread per line of p-1 entries at stride R
template <class R>
__device__ __forceinline__ void pass_line(const R& ring, const Pass& ps, typename R::T* cur, int n)
{
  typedef typename R::T T;
  const int p = ps.p, len = ps.p - 1, Rr = ps.R;
  const int lines = n / len;
  for (int ln = threadIdx.x; ln < lines; ln += blockDim.x) {
    int r = ln % Rr, blk = ln / Rr;
    T* v = cur + (size_t)blk * len * Rr + r;
#define V(a) v[(size_t)(a) * Rr]
    switch (ps.kind) {
      case PASS_L:
        for (int a = 1; a < len; a++) V(a) = ring.add(V(a), V(a - 1));
        break;
      case PASS_LINV:
        for (int a = len - 1; a >= 1; a--) V(a) = ring.sub(V(a), V(a - 1));
        break;
      case PASS_GPOW: {
        T last = V(len - 1);
        for (int a = len - 1; a >= 1; a--) V(a) = ring.add(V(a), ring.sub(last, V(a - 1)));
        V(0) = ring.add(V(0), last);
      } break;
      case PASS_GDEC: {
        T acc = V(0);
        for (int a = len - 1; a >= 1; a--) {
          acc = ring.add(acc, V(a));
          V(a) = ring.sub(V(a), V(a - 1));
        }
        V(0) = ring.add(V(0), acc);
      } break;
      case PASS_GINVPOW: {
        T lo = ring.zero(), hi = ring.zero();
        for (int a = 0; a < len; a++) lo = ring.add(lo, V(a));
        for (int a = len - 1; a >= 0; a--) {
          T z = V(a);
          V(a) = ring.sub(ring.mul(ring.from_int(p - 1 - a), lo), ring.mul(ring.from_int(a + 1), hi));
          lo = ring.sub(lo, z);
          hi = ring.add(hi, z);
        }
      } break;
      case PASS_GINVDEC: {
        T s = ring.zero();
        for (int a = 0; a < len; a++) s = ring.add(s, ring.mul(ring.from_int(a + 1), V(a)));
        T acc = s;
        T pp = ring.from_int(p);
        for (int a = len - 1; a >= 1; a--) {
          T keep = acc;
          acc = ring.sub(acc, ring.mul(V(a), pp));
          V(a) = keep;
        }
        V(0) = acc;
      } break;
      case PASS_NORMSQ: {
        T s = ring.zero();
        for (int a = 0; a < len; a++) s = ring.add(s, V(a));
        for (int a = 0; a < len; a++) V(a) = ring.add(V(a), s);
      } break;
      default: break;
    }
#undef V
  }
}

// random.cpp:19-50: out[row] = (sum_{col=1}^{p-1} 2*c(row,col)*in[col-1]) / sqrt(2), cur -> alt
__device__ __forceinline__ void pass_gauss(const Pass& ps, const double* cur, double* alt, int n, const double2* tab)
{
  const int p = ps.p, d = ps.d, Rr = ps.R, rs = ps.rustride, half = ps.p >> 1;
  const double2* ru = tab + ps.tab;
  const double sqrt2 = sqrt(2.0);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int r = j % Rr, t = j / Rr;
    int row = t % d, blk = t / d;
    const double* in = cur + (size_t)blk * d * Rr + r;
    double acc = 0.0;
    int idx = row;                      // (row*col) mod p at col = 1
    for (int col = 1; col <= p - 1; col++) {
      double2 w = ru[(size_t)idx * rs];
      double c = col <= half ? w.x : w.y;
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(2.0, c), in[(size_t)(col - 1) * Rr]));
      idx += row; if (idx >= p) idx -= p;
    }
    alt[j] = __ddiv_rn(acc, sqrt2);
  }
}

// ------------------------------------------------------------------ kernels

template <class R, class W>
struct CrtParams {
  typename R::IO* y;
  int64_t batch;
  int32_t n, k;
  const W* tab;
  int32_t tab_stride;
  typename R::T* ws;          // nullptr: shared memory
  int32_t finish;             // FIN_NONE / FIN_SCALE
  ZqConsts zc;                // Zq only
  double2 cscale[kMaxLimbs];  // C64 only
};

__device__ __forceinline__ ZqRing make_ring(const ZqRing*, const ZqConsts& zc, int limb) { return ZqRing::make(zc, limb); }
__device__ __forceinline__ C64Ring make_ring(const C64Ring*, const ZqConsts&, int) { return C64Ring{}; }

template <class R, class W>
__global__ void __launch_bounds__(kEngineThreads)
k_engine_crt(const __grid_constant__ CrtParams<R, W> P, const __grid_constant__ PassList PL)
{
  typedef typename R::T T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = P.n, k = P.k;
  T* buf0 = P.ws ? P.ws + (size_t)blockIdx.x * 2 * n : reinterpret_cast<T*>(smem_raw);
  T* buf1 = buf0 + n;
  const int64_t work_items = P.batch * k;
  for (int64_t work = blockIdx.x; work < work_items; work += gridDim.x) {
    const int64_t e = work / k;
    const int limb = (int)(work - e * k);
    const R ring = make_ring((const R*)nullptr, P.zc, limb);
    typename R::IO* base = P.y + (size_t)e * n * k + limb;
    const W* tab = P.tab + (size_t)limb * P.tab_stride;
    T* cur = buf0;
    T* alt = buf1;
    for (int j = threadIdx.x; j < n; j += blockDim.x) cur[j] = ring.load(base[(size_t)j * k]);
    __syncthreads();
    for (int i = 0; i < PL.count; i++) {
      const Pass& ps = PL.pass[i];
      if (ps.kind == PASS_DIAG) {
        pass_diag(ring, ps, cur, n, tab);
      } else if (ps.kind == PASS_DFT && ps.p == 2) {
        pass_dft2(ring, ps, cur, n);
      } else {
        pass_dense(ring, ps, cur, alt, n, tab);
        T* tmp = cur; cur = alt; alt = tmp;
      }
      __syncthreads();
    }
    if (P.finish == FIN_SCALE) {
      if constexpr (sizeof(T) == 4) {
        const T s = P.zc.scale[limb];
        for (int j = threadIdx.x; j < n; j += blockDim.x) base[(size_t)j * k] = ring.store(ring.mul(cur[j], s));
      } else {
        const T s = P.cscale[limb];
        for (int j = threadIdx.x; j < n; j += blockDim.x) base[(size_t)j * k] = ring.store(ring.mul(cur[j], s));
      }
    } else {
      for (int j = threadIdx.x; j < n; j += blockDim.x) base[(size_t)j * k] = ring.store(cur[j]);
    }
    __syncthreads();
  }
}

// line engine.  per_limb = 1: CTA per (element, limb), stride k (Zq).  per_limb = 0: CTA per element with the
// limbs folded into the right stride (modulus-free rings: I (x) A (x) I_{R*k}), n = totm*k, stride 1.
template <class R>
struct LineParams {
  typename R::IO* y;
  int64_t batch;
  int32_t n, k;               // n already folded when per_limb == 0 (then k == 1 here)
  typename R::T* ws;
  int32_t finish;
  ZqConsts zc;
  int64_t divisor;            // FIN_DIV_EXACT
  int16_t* ok;                // FIN_DIV_EXACT
  double rscale;              // FIN_REAL_SCALE
};

__device__ __forceinline__ ZqRing make_line_ring(const ZqRing*, const ZqConsts& zc, int limb) { return ZqRing::make(zc, limb); }
__device__ __forceinline__ I64Ring make_line_ring(const I64Ring*, const ZqConsts&, int) { return I64Ring{}; }
__device__ __forceinline__ F64Ring make_line_ring(const F64Ring*, const ZqConsts&, int) { return F64Ring{}; }
__device__ __forceinline__ C64Ring make_line_ring(const C64Ring*, const ZqConsts&, int) { return C64Ring{}; }

template <class R> struct RingTag { static constexpr int id = -1; };
template <> struct RingTag<ZqRing> { static constexpr int id = RING_ZQ; };
template <> struct RingTag<I64Ring> { static constexpr int id = RING_I64; };
template <> struct RingTag<F64Ring> { static constexpr int id = RING_F64; };
template <> struct RingTag<C64Ring> { static constexpr int id = RING_C64; };

template <class R>
__global__ void __launch_bounds__(kEngineThreads)
k_engine_line(const __grid_constant__ LineParams<R> P, const __grid_constant__ PassList PL)
{
  typedef typename R::T T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = P.n, k = P.k;
  T* cur = P.ws ? P.ws + (size_t)blockIdx.x * n : reinterpret_cast<T*>(smem_raw);
  const int64_t work_items = P.batch * k;
  for (int64_t work = blockIdx.x; work < work_items; work += gridDim.x) {
    const int64_t e = work / k;
    const int limb = (int)(work - e * k);
    const R ring = make_line_ring((const R*)nullptr, P.zc, limb);
    typename R::IO* base = P.y + (size_t)e * n * k + limb;
    for (int j = threadIdx.x; j < n; j += blockDim.x) cur[j] = ring.load(base[(size_t)j * k]);
    __syncthreads();
    for (int i = 0; i < PL.count; i++) {
      pass_line(ring, PL.pass[i], cur, n);
      __syncthreads();
    }
    if constexpr (RingTag<R>::id == RING_ZQ) {
      if (P.finish == FIN_SCALE) {
        const T s = P.zc.scale[limb];
        for (int j = threadIdx.x; j < n; j += blockDim.x) base[(size_t)j * k] = ring.store(ring.mul(cur[j], s));
      } else {
        for (int j = threadIdx.x; j < n; j += blockDim.x) base[(size_t)j * k] = ring.store(cur[j]);
      }
    } else if constexpr (RingTag<R>::id == RING_I64) {
      if (P.finish == FIN_DIV_EXACT) {
        int bad = 0;
        for (int j = threadIdx.x; j < n; j += blockDim.x) bad |= (cur[j] % P.divisor) != 0;
        const int all_ok = __syncthreads_and(!bad);
        for (int j = threadIdx.x; j < n; j += blockDim.x) base[j] = all_ok ? cur[j] / P.divisor : cur[j];
        if (threadIdx.x == 0 && P.ok) P.ok[e] = (int16_t)all_ok;
      } else {
        for (int j = threadIdx.x; j < n; j += blockDim.x) base[j] = cur[j];
      }
    } else if constexpr (RingTag<R>::id == RING_C64) {
      if (P.finish == FIN_REAL_SCALE) {
        for (int j = threadIdx.x; j < n; j += blockDim.x)
          base[j] = make_double2(__dmul_rn(cur[j].x, P.rscale), __dmul_rn(cur[j].y, P.rscale));
      } else {
        for (int j = threadIdx.x; j < n; j += blockDim.x) base[j] = cur[j];
      }
    } else {
      for (int j = threadIdx.x; j < n; j += blockDim.x) base[j] = cur[j];
    }
    __syncthreads();
  }
}

struct GaussParams {
  double* y;
  int64_t batch;
  int32_t n, k;
  const double2* tab;
  int32_t tab_stride;
  double* ws;
};

__global__ void __launch_bounds__(kEngineThreads)
k_engine_gauss(const __grid_constant__ GaussParams P, const __grid_constant__ PassList PL)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = P.n, k = P.k;
  double* buf0 = P.ws ? P.ws + (size_t)blockIdx.x * 2 * n : reinterpret_cast<double*>(smem_raw);
  double* buf1 = buf0 + n;
  const int64_t work_items = P.batch * k;
  for (int64_t work = blockIdx.x; work < work_items; work += gridDim.x) {
    const int64_t e = work / k;
    const int limb = (int)(work - e * k);
    double* base = P.y + (size_t)e * n * k + limb;
    const double2* tab = P.tab + (size_t)limb * P.tab_stride;
    double* cur = buf0;
    double* alt = buf1;
    for (int j = threadIdx.x; j < n; j += blockDim.x) cur[j] = base[(size_t)j * k];
    __syncthreads();
    for (int i = 0; i < PL.count; i++) {
      pass_gauss(PL.pass[i], cur, alt, n, tab);
      double* tmp = cur; cur = alt; alt = tmp;
      __syncthreads();
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) base[(size_t)j * k] = cur[j];
    __syncthreads();
  }
}

// norm.cpp:39-80: out[e*k+limb] = sum_j y[j] * ((x)(I+J) y)[j]
template <class R>
struct NormParams {
  const typename R::IO* y;
  typename R::IO* out;
  int64_t batch;
  int32_t n, k;
  typename R::T* ws;
};

template <class T> __device__ __forceinline__ T warp_sum(T v, const I64Ring&)
{
  for (int o = 16; o > 0; o >>= 1) v = (T)((uint64_t)v + (uint64_t)__shfl_down_sync(0xffffffffu, v, o));
  return v;
}
template <class T> __device__ __forceinline__ T warp_sum(T v, const F64Ring&)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

template <class R>
__global__ void __launch_bounds__(kEngineThreads)
k_engine_normsq(const __grid_constant__ NormParams<R> P, const __grid_constant__ PassList PL)
{
  typedef typename R::T T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ T red[32];
  const int n = P.n, k = P.k;
  T* cur = P.ws ? P.ws + (size_t)blockIdx.x * 2 * n : reinterpret_cast<T*>(smem_raw);
  T* orig = cur + n;
  const R ring{};
  const int64_t work_items = P.batch * k;
  for (int64_t work = blockIdx.x; work < work_items; work += gridDim.x) {
    const int64_t e = work / k;
    const int limb = (int)(work - e * k);
    const typename R::IO* base = P.y + (size_t)e * n * k + limb;
    for (int j = threadIdx.x; j < n; j += blockDim.x) { T v = base[(size_t)j * k]; cur[j] = v; orig[j] = v; }
    __syncthreads();
    for (int i = 0; i < PL.count; i++) {
      pass_line(ring, PL.pass[i], cur, n);
      __syncthreads();
    }
    T acc = ring.zero();
    for (int j = threadIdx.x; j < n; j += blockDim.x) acc = ring.add(acc, ring.mul(orig[j], cur[j]));
    acc = warp_sum(acc, ring);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      T v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : ring.zero();
      v = warp_sum(v, ring);
      if (threadIdx.x == 0) P.out[e * k + limb] = v;
    }
    __syncthreads();
  }
}

// mul.cpp:14-35, thread per coefficient
__global__ void k_mul_zq(int64_t* __restrict__ a, const int64_t* __restrict__ b, int64_t total, int64_t nk, int k,
                         int broadcast, const __grid_constant__ ZqConsts zc)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int limb = (int)(i % k);
    const ZqRing ring = ZqRing::make(zc, limb);
    const int64_t bi = broadcast ? i % nk : i;
    a[i] = ring.store(ring.mul(ring.load(a[i]), ring.load(b[bi])));
  }
}

__global__ void k_mul_c(double2* __restrict__ a, const double2* __restrict__ b, int64_t total, int64_t nk, int broadcast)
{
  const C64Ring ring{};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bi = broadcast ? i % nk : i;
    a[i] = ring.mul(a[i], b[bi]);
  }
}

// ------------------------------------------------------------------ launchers

struct LaunchShape {
  int grid;
  size_t smem;
  bool use_ws;
  void* ws;      // the stream's workspace when use_ws
};

// buffers: how many n-element buffers the kernel keeps; elem_bytes: sizeof(T)
static int pick_shape(const lolb_plan* pl, const void* kernel, size_t bytes_per_cta, int64_t work_items, cudaStream_t st,
                      LaunchShape* out, size_t ws_bytes_per_cta = 0)
{
  out->use_ws = bytes_per_cta > kSmemBudget;
  out->ws = nullptr;
  out->smem = out->use_ws ? 0 : bytes_per_cta;
  int per_sm = 8;
  if (!out->use_ws && bytes_per_cta > 0) {
    per_sm = (int)((220 * 1024) / (bytes_per_cta + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
  }
  int64_t g = (int64_t)pl->num_sms * per_sm;
  if (g > work_items) g = work_items;
  if (g < 1) g = 1;
  out->grid = (int)g;
  if (!out->use_ws && out->smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)out->smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  }
  if (out->use_ws) {
    out->ws = plan_ws(pl, st, (size_t)out->grid * (ws_bytes_per_cta ? ws_bytes_per_cta : bytes_per_cta));
    if (!out->ws) return LOLB_ERR_CUDA;
  }
  return LOLB_OK;
}

static int check_launch(const char* what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  count_launch();
  return LOLB_OK;
}

int engine_crt_zq(const lolb_plan* pl, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  const PassList& PL = inverse ? pl->crt_inv : pl->crt_fwd;
  CrtParams<ZqRing, uint32_t> P{};
  P.y = y; P.batch = batch; P.n = pl->n; P.k = pl->k;
  P.tab = inverse ? pl->d_tab_inv : pl->d_tab_fwd;
  P.tab_stride = inverse ? pl->tab_stride_inv : pl->tab_stride_fwd;
  P.finish = inverse ? FIN_SCALE : FIN_NONE;
  P.zc = inverse ? pl->zq_mhat : pl->zq_plain;
  LaunchShape sh;
  const void* kern = (const void*)k_engine_crt<ZqRing, uint32_t>;
  // the kernel always addresses two buffers when it uses the workspace
  int rc = pick_shape(pl, kern, (size_t)pl->n * 4 * (PL.needs_alt ? 2 : 1) * 1, batch * pl->k, st, &sh, (size_t)2 * pl->n * 4);
  if (rc) return rc;
  P.ws = (uint32_t*)sh.ws;
  k_engine_crt<ZqRing, uint32_t><<<sh.grid, kEngineThreads, sh.smem, st>>>(P, PL);
  return check_launch("k_engine_crt<Zq>");
}

int engine_crt_c(const lolb_plan* pl, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  const PassList& PL = inverse ? pl->crt_inv : pl->crt_fwd;
  CrtParams<C64Ring, double2> P{};
  P.y = y; P.batch = batch; P.n = pl->n; P.k = pl->k;
  P.tab = inverse ? pl->d_ctab_inv : pl->d_ctab_fwd;
  P.tab_stride = inverse ? pl->ctab_stride_inv : pl->ctab_stride_fwd;
  P.finish = inverse ? FIN_SCALE : FIN_NONE;
  for (int i = 0; i < pl->k; i++) P.cscale[i] = pl->c_mhatinv[i];
  LaunchShape sh;
  const void* kern = (const void*)k_engine_crt<C64Ring, double2>;
  int rc = pick_shape(pl, kern, (size_t)pl->n * 16 * (PL.needs_alt ? 2 : 1), batch * pl->k, st, &sh, (size_t)2 * pl->n * 16);
  if (rc) return rc;
  P.ws = (double2*)sh.ws;
  k_engine_crt<C64Ring, double2><<<sh.grid, kEngineThreads, sh.smem, st>>>(P, PL);
  return check_launch("k_engine_crt<C64>");
}

template <class R>
static int launch_line(const lolb_plan* pl, LineParams<R>& P, const PassList& PL, cudaStream_t st, const char* what)
{
  LaunchShape sh;
  const void* kern = (const void*)k_engine_line<R>;
  int rc = pick_shape(pl, kern, (size_t)P.n * sizeof(typename R::T), P.batch * P.k, st, &sh);
  if (rc) return rc;
  P.ws = (typename R::T*)sh.ws;
  k_engine_line<R><<<sh.grid, kEngineThreads, sh.smem, st>>>(P, PL);
  return check_launch(what);
}

int engine_line_zq(const lolb_plan* pl, int kind, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  LineParams<ZqRing> P{};
  P.y = y; P.batch = batch; P.n = pl->n; P.k = pl->k; P.zc = zc;
  P.finish = scale ? FIN_SCALE : FIN_NONE;
  return launch_line<ZqRing>(pl, P, pl->line[kind], st, "k_engine_line<Zq>");
}

int engine_line_i64(const lolb_plan* pl, int kind, int64_t divisor, int16_t* ok, int64_t* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  LineParams<I64Ring> P{};
  P.y = y; P.batch = batch; P.n = pl->n * pl->k; P.k = 1;
  P.finish = divisor ? FIN_DIV_EXACT : FIN_NONE;
  P.divisor = divisor; P.ok = ok;
  return launch_line<I64Ring>(pl, P, pl->line_folded[kind], st, "k_engine_line<I64>");
}

int engine_line_f64(const lolb_plan* pl, int kind, double* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  LineParams<F64Ring> P{};
  P.y = y; P.batch = batch; P.n = pl->n * pl->k; P.k = 1;
  return launch_line<F64Ring>(pl, P, pl->line_folded[kind], st, "k_engine_line<F64>");
}

int engine_line_c64(const lolb_plan* pl, int kind, double rscale, double2* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  LineParams<C64Ring> P{};
  P.y = y; P.batch = batch; P.n = pl->n * pl->k; P.k = 1;
  P.finish = rscale != 0.0 ? FIN_REAL_SCALE : FIN_NONE;
  P.rscale = rscale;
  return launch_line<C64Ring>(pl, P, pl->line_folded[kind], st, "k_engine_line<C64>");
}

int engine_gauss(const lolb_plan* pl, double* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  GaussParams P{};
  P.y = y; P.batch = batch; P.n = pl->n; P.k = pl->k;
  P.tab = pl->d_ctab_fwd; P.tab_stride = pl->ctab_stride_fwd;
  LaunchShape sh;
  int rc = pick_shape(pl, (const void*)k_engine_gauss, (size_t)pl->n * 8 * 2, batch * pl->k, st, &sh);
  if (rc) return rc;
  P.ws = (double*)sh.ws;
  k_engine_gauss<<<sh.grid, kEngineThreads, sh.smem, st>>>(P, pl->line[PASS_GAUSS]);
  return check_launch("k_engine_gauss");
}

template <class R>
static int launch_normsq(const lolb_plan* pl, const typename R::IO* y, typename R::IO* out, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  NormParams<R> P{};
  P.y = y; P.out = out; P.batch = batch; P.n = pl->n; P.k = pl->k;
  LaunchShape sh;
  int rc = pick_shape(pl, (const void*)k_engine_normsq<R>, (size_t)pl->n * sizeof(typename R::T) * 2, batch * pl->k, st, &sh);
  if (rc) return rc;
  P.ws = (typename R::T*)sh.ws;
  k_engine_normsq<R><<<sh.grid, kEngineThreads, sh.smem, st>>>(P, pl->line[PASS_NORMSQ]);
  return check_launch("k_engine_normsq");
}

int engine_normsq_i64(const lolb_plan* pl, const int64_t* y, int64_t* out, int64_t batch, cudaStream_t st)
{ return launch_normsq<I64Ring>(pl, y, out, batch, st); }
int engine_normsq_f64(const lolb_plan* pl, const double* y, double* out, int64_t batch, cudaStream_t st)
{ return launch_normsq<F64Ring>(pl, y, out, batch, st); }

int engine_mul_zq(const lolb_plan* pl, int64_t* a, const int64_t* b, int64_t batch, int64_t b_batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  const int64_t nk = (int64_t)pl->n * pl->k, total = batch * nk;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)pl->num_sms * 16) blocks = (int64_t)pl->num_sms * 16;
  k_mul_zq<<<(int)blocks, 256, 0, st>>>(a, b, total, nk, pl->k, b_batch == 1 ? 1 : 0, pl->zq_plain);
  return check_launch("k_mul_zq");
}

int engine_mul_c(const lolb_plan* pl, double2* a, const double2* b, int64_t batch, int64_t b_batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  const int64_t nk = (int64_t)pl->n * pl->k, total = batch * nk;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)pl->num_sms * 16) blocks = (int64_t)pl->num_sms * 16;
  k_mul_c<<<(int)blocks, 256, 0, st>>>(a, b, total, nk, b_batch == 1 ? 1 : 0);
  return check_launch("k_mul_c");
}

}  // namespace lolb
