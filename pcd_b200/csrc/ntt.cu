// ntt.cu -- radix-2 (coset) NTT over the two 298-bit scalar fields of the MNT cycle.
//
// Replaces ark-poly Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place as driven by
// ark-groth16's R1CStoQAP::witness_map (reached from
// /root/reference/src/ec_cycle_pcd/mod.rs:171,179).  Natural order in, natural order out,
// out[i] = sum_j in[j] * omega^(i j) with omega_n = TWO_ADIC_ROOT^(2^(s - log n)).
//
// Algorithm: decimation in time over a bit-reversed load.  The log n butterfly stages are cut
// into passes of r <= 8 stages; one CTA owns a tile of 2^r "rows" (the indices that interact in
// those stages) x W "columns" (neighbouring independent transforms, so that every global access is
// a run of W contiguous 40-byte elements), keeps it in shared memory in limb-major planes
// (swizzled so that the butterflies are bank-conflict free, ntt_sm_index) and runs the r stages there.  The first pass folds the
// bit reversal into its (strided) load and writes contiguous rows; later passes work in place.
// Coset scaling (g^i before a forward transform, g^-i / n after an inverse one) is fused into
// the first load / last store.  HBM traffic per pass = one read + one write of the vector.
#include "common.cuh"
#include "ntt.cuh"

#include <cstdlib>
static constexpr int NTT_THREADS = 256;  // upper bound (launch bounds); the launch uses ntt_threads()
static constexpr int NTT_MAX_R = 8;
// Elements per CTA tile and threads per CTA.  Measured on B200 (tools/probe_ntt.py, coset FFT over r4, ms at
// 2^16 / 2^20 / 2^24): tile 2^11 x 256 threads 0.133 / 0.534 / 8.29;  2^10 x 128: 0.096 / 0.436 / 7.42;
// 2^9 x 128: 0.054 / 0.416 / 7.29 (chosen);  2^9 x 256: 0.045 / 0.507 / 9.27;  2^8 x 64: 0.056 / 0.412 / 7.94.
// A 20 KiB tile lets ~10 CTAs share an SM, so one CTA's barriers and twiddle loads hide behind the others'
// butterflies; the shorter runs of contiguous elements per row (2^9 / 2^r x 40 B) cost nothing because DRAM is
// < 10 % busy.  PCDGPU_NTT_TILE_LOG / PCDGPU_NTT_THREADS override them for sweeps (development aid).
static int ntt_tile_log() {
  static const int v = getenv("PCDGPU_NTT_TILE_LOG") ? atoi(getenv("PCDGPU_NTT_TILE_LOG")) : 9;
  return v < NTT_MAX_R ? NTT_MAX_R : (v > 12 ? 12 : v);
}
static int ntt_threads() {
  static const int v = getenv("PCDGPU_NTT_THREADS") ? atoi(getenv("PCDGPU_NTT_THREADS")) : 128;
  return v < 32 ? 32 : (v > NTT_THREADS ? NTT_THREADS : v);
}

template <class F>
__device__ __forceinline__ F ld_elem(const u32* g, size_t idx) {
  const uint2* p = reinterpret_cast<const uint2*>(g + idx * 10);
  F r;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint2 v = p[i];
    r.l[2 * i] = v.x;
    r.l[2 * i + 1] = v.y;
  }
  return r;
}
template <class F>
__device__ __forceinline__ F ldg_elem(const u32* g, size_t idx) {
  const uint2* p = reinterpret_cast<const uint2*>(g + idx * 10);
  F r;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint2 v = __ldg(p + i);
    r.l[2 * i] = v.x;
    r.l[2 * i + 1] = v.y;
  }
  return r;
}
template <class F>
__device__ __forceinline__ void st_elem(u32* g, size_t idx, const F& a) {
  uint2* p = reinterpret_cast<uint2*>(g + idx * 10);
#pragma unroll
  for (int i = 0; i < 5; i++) p[i] = make_uint2(a.l[2 * i], a.l[2 * i + 1]);
}

struct NttPass {
  const u32* src;
  u32* dst;
  const u32* tw;    // omega^k, k < n/2
  const u32* pre;   // first pass: multiply input i by pre[i] (or null)
  const u32* post;  // last pass: multiply output i by post[i] (or null)
  const u32* post_c;  // last pass: multiply every output by *post_c (or null)
  int log_n, s, r, logW;
  int inverse, first, last;
  size_t batch_stride;  // elements between consecutive transforms of a batch (blockIdx.y)
};

// Shared-memory word of element (row j, column c) inside a limb plane: i = j W + c with the low five bits flipped
// when bit 5 is set.  In butterfly stage u the 32 lanes of a warp touch rows whose index has bit u fixed, i.e.
// i has one of its low five bits fixed and bit 5 varying -- with a linear (or padded-pitch) layout that is a
// two-way bank conflict on every access of the first 5 - log W stages (ncu: 21 M conflicts per pass); the flip
// moves the two halves onto complementary banks.  Conflict free for every stage, W and r (checked by enumeration).
__device__ __forceinline__ int ntt_sm_index(int j, int c, int logW) {
  int i = (j << logW) | c;
  return i ^ ((i & 32) ? 31 : 0);
}

template <class F>
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(NttPass a) {
  extern __shared__ u32 sm[];
  const int W = 1 << a.logW;
  const int R = 1 << a.r;
  const int tile = 1 << (a.r + a.logW);
  const int plane = tile;             // words per limb plane
  (void)R;
  const int k = a.log_n;
  const size_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  a.src += (size_t)blockIdx.y * a.batch_stride * 10;
  a.dst += (size_t)blockIdx.y * a.batch_stride * 10;

  size_t lowbase = 0, high = 0;
  if (!a.first) {
    int lb = a.s - a.logW;  // tiles per "low" range
    lowbase = (b & (((size_t)1 << lb) - 1)) << a.logW;
    high = b >> lb;
  }
  // ---- load ------------------------------------------------------------------------------
  for (int e = tid; e < tile; e += nthr) {
    int j = e >> a.logW, c = e & (W - 1);
    size_t idx;
    if (a.first) {
      size_t jr = __brev((unsigned)j) >> (32 - a.r);
      if (a.r == 0) jr = 0;
      idx = (jr << (k - a.r)) | ((b << a.logW) | (size_t)c);
    } else {
      idx = (high << (a.s + a.r)) | ((size_t)j << a.s) | lowbase | (size_t)c;
    }
    F v = ld_elem<F>(a.src, idx);
    if (a.first && a.pre) v = v * ldg_elem<F>(a.pre, idx);
    u32* dst = sm + ntt_sm_index(j, c, a.logW);
#pragma unroll
    for (int l = 0; l < 10; l++) dst[l * plane] = v.l[l];
  }
  // ---- r butterfly stages in shared memory -------------------------------------------------
  const size_t half_n = (size_t)1 << (k - 1);
  for (int u = 0; u < a.r; u++) {
    __syncthreads();
    const int sh = k - a.s - u - 1;
    for (int q = tid; q < (tile >> 1); q += nthr) {
      int c = q & (W - 1), jj = q >> a.logW;
      int lo = jj & ((1 << u) - 1);
      int j = ((jj >> u) << (u + 1)) | lo;
      u32* p1 = sm + ntt_sm_index(j, c, a.logW);
      u32* p2 = sm + ntt_sm_index(j + (1 << u), c, a.logW);
      F A, B;
#pragma unroll
      for (int l = 0; l < 10; l++) {
        A.l[l] = p1[l * plane];
        B.l[l] = p2[l * plane];
      }
      size_t te = (((size_t)lo << a.s) + (a.first ? 0 : (lowbase + (size_t)c))) << sh;
      F X, Y;
      if (te == 0) {
        X = A + B;
        Y = A - B;
      } else if (!a.inverse) {
        F t = B * ldg_elem<F>(a.tw, te);
        X = A + t;
        Y = A - t;
      } else {
        F t = B * ldg_elem<F>(a.tw, half_n - te);  // omega^-e = -omega^(n/2 - e)
        X = A - t;
        Y = A + t;
      }
#pragma unroll
      for (int l = 0; l < 10; l++) {
        p1[l * plane] = X.l[l];
        p2[l * plane] = Y.l[l];
      }
    }
  }
  __syncthreads();
  // ---- store -----------------------------------------------------------------------------
  F pc;
  if (a.last && a.post_c) pc = ldg_elem<F>(a.post_c, 0);
  for (int e = tid; e < tile; e += nthr) {
    int j, c;
    size_t idx;
    if (a.first) {  // rows are contiguous in the output: j fastest
      j = e & (R - 1);
      c = e >> a.r;
      size_t col = (b << a.logW) | (size_t)c;
      int hb = k - a.r;  // bit-reverse the column index over k - r bits
      size_t cr = 0;
      if (hb > 0) cr = (size_t)(__brevll((unsigned long long)col) >> (64 - hb));
      idx = (cr << a.r) | (size_t)j;
    } else {
      j = e >> a.logW;
      c = e & (W - 1);
      idx = (high << (a.s + a.r)) | ((size_t)j << a.s) | lowbase | (size_t)c;
    }
    const u32* srcp = sm + ntt_sm_index(j, c, a.logW);
    F v;
#pragma unroll
    for (int l = 0; l < 10; l++) v.l[l] = srcp[l * plane];
    if (a.last) {
      if (a.post) v = v * ldg_elem<F>(a.post, idx);
      else if (a.post_c) v = v * pc;
    }
    st_elem<F>(a.dst, idx, v);
  }
}

// tw[k] = omega^k (k < n/2); cpow[i] = g^i; cinv[i] = g^-i / n; consts = {1/n, 1/(g^n - 1)}
template <class F>
__global__ void ntt_tables_kernel(u32* tw, u32* cpow, u32* cinv, u32* consts, int log_n) {
  typedef typename F::Params P;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t n = (size_t)1 << log_n;
  if (i >= n) return;
  F g, gi, ti, w;
#pragma unroll
  for (int l = 0; l < 10; l++) {
    g.l[l] = P::generator(l);
    gi.l[l] = P::generator_inv(l);
    ti.l[l] = P::two_inv(l);
    w.l[l] = P::two_adic_root(l);
  }
  F ninv = F::one();
  for (int b = 0; b < log_n; b++) ninv = ninv * ti;
  st_elem<F>(cpow, i, g.pow64(i));
  st_elem<F>(cinv, i, gi.pow64(i) * ninv);
  if (i < n / 2) {
    for (int b = log_n; b < P::TWO_ADICITY; b++) w = w.sqr();
    st_elem<F>(tw, i, w.pow64(i));
  }
  if (i == 0) {
    st_elem<F>(consts, 0, ninv);
    F gn = g.pow64(n) - F::one();
    st_elem<F>(consts, 1, gn.inverse());
  }
}

template <class F>
static int ntt_tables_get(pcdgpu_ctx* ctx, int field, int log_n, NttTablesDev* out) {
  int key = field * 64 + log_n;
  auto it = ctx->ntt_tables.find(key);
  if (it == ctx->ntt_tables.end()) {
    size_t n = (size_t)1 << log_n;
    NttTables t;
    void* all = nullptr;
    size_t half = n / 2 ? n / 2 : 1;
    size_t bytes = (half + 2 * n + 2) * 40;
    cudaError_t e = cudaMalloc(&all, bytes);
    if (e != cudaSuccess) {
      ctx->set_error("cudaMalloc(%zu) for NTT tables: %s", bytes, cudaGetErrorString(e));
      return PCDGPU_E_NOMEM;
    }
    t.twiddles = all;
    t.coset_pow = (char*)all + half * 40;
    t.coset_inv = (char*)t.coset_pow + n * 40;
    u32* consts = (u32*)((char*)t.coset_inv + n * 40);
    unsigned grid = (unsigned)((n + 127) / 128);
    ntt_tables_kernel<F><<<grid, 128, 0, ctx->stream>>>((u32*)t.twiddles, (u32*)t.coset_pow, (u32*)t.coset_inv,
                                                       consts, log_n);
    PCD_CUDA(ctx, cudaGetLastError());
    it = ctx->ntt_tables.emplace(key, t).first;
  }
  size_t n = (size_t)1 << log_n;
  out->tw = (const u32*)it->second.twiddles;
  out->cpow = (const u32*)it->second.coset_pow;
  out->cinv = (const u32*)it->second.coset_inv;
  out->ninv = (const u32*)((const char*)it->second.coset_inv + n * 40);
  out->zinv = out->ninv + 10;
  return 0;
}

int ntt_tables(pcdgpu_ctx* ctx, int field, int log_n, NttTablesDev* out) {
  if (field == PCDGPU_FIELD_R4) return ntt_tables_get<FpR4>(ctx, field, log_n, out);
  return ntt_tables_get<FpQ4>(ctx, field, log_n, out);
}

template <class F>
static int ntt_run_t(pcdgpu_ctx* ctx, int field, void* d_data, int log_n, int inverse, int coset, int batch = 1) {
  if (log_n > F::Params::TWO_ADICITY) {
    ctx->set_error("radix-2 domain 2^%d exceeds the field's 2-adicity %d", log_n, F::Params::TWO_ADICITY);
    return PCDGPU_E_DOMAIN;
  }
  if (log_n == 0) return 0;  // n = 1: every flavour is the identity
  NttTablesDev t;
  PCD_TRY(ntt_tables(ctx, field, log_n, &t));
  const int NTT_TILE_LOG = ntt_tile_log();
  // largest tile in shared memory: a multi-pass tile (2^tile elements) or the single-pass case (2^10)
  int max_smem = 10 * (1 << NTT_TILE_LOG) * 4;
  if (max_smem < 10 * 1024 * 4) max_smem = 10 * 1024 * 4;
  PCD_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  int passes = log_n <= 10 ? 1 : (log_n + NTT_MAX_R - 1) / NTT_MAX_R;
  void* scratch = nullptr;
  if (passes > 1) PCD_TRY(ctx->scratch(SLOT_NTT, (((size_t)40) << log_n) * batch, &scratch));
  int s = 0;
  int ps = ctx->prof_begin(PROF_NTT, (double)batch * (double)log_n * (double)((size_t)1 << (log_n - 1)));
  ctx->launches += passes;
  for (int p = 0; p < passes; p++) {
    int r = (log_n - s + (passes - p) - 1) / (passes - p);  // spread stages evenly, larger first
    NttPass a;
    a.first = p == 0;
    a.last = p == passes - 1;
    a.src = (const u32*)(p == 0 ? d_data : scratch);
    a.dst = (u32*)((a.last || passes == 1) ? d_data : scratch);
    a.tw = t.tw;
    a.pre = (a.first && coset && !inverse) ? t.cpow : nullptr;
    a.post = (a.last && coset && inverse) ? t.cinv : nullptr;
    a.post_c = (a.last && inverse && !coset) ? t.ninv : nullptr;
    a.log_n = log_n;
    a.s = s;
    a.r = r;
    a.logW = passes == 1 ? 0 : NTT_TILE_LOG - r;
    a.inverse = inverse;
    a.batch_stride = (size_t)1 << log_n;
    int R = 1 << r, W = 1 << a.logW;
    size_t smem = (size_t)10 * R * W * 4;
    dim3 grid((unsigned)(((size_t)1 << log_n) >> (r + a.logW)), (unsigned)batch);
    ntt_pass_kernel<F><<<grid, ntt_threads(), smem, ctx->stream>>>(a);
    PCD_CUDA(ctx, cudaGetLastError());
    s += r;
  }
  ctx->prof_end(ps);
  return 0;
}

// ---- mixed-radix domains 7^a 2^b (q4 only) ----------------------------------------------------------
// ark-poly MixedRadixEvaluationDomain, used by GeneralEvaluationDomain::new when a size exceeds 2^17 on
// q4 = MNT6-298 Fr.  Cooley-Tukey: `a` split levels peel the factors of 7
//   y[k1][n2] = w_L^(n2 k1) * sum_{n1 < 7} x[M n1 + n2] w_7^(n1 k1),   L = 7 M,
// leaving 7^a contiguous rows of length 2^b for the batched radix-2 kernel; a final gather writes
// X[k1 + 7 k1' + 49 k2] in natural order and applies the inverse / coset scaling.
// wpow[j] = w_N^j (j < N); cpow[i] = g^i; cinv[i] = g^-i / 7^a; consts[0] = 1 / 7^a
template <class F>
__global__ void mixed_tables_kernel(u32* wpow, u32* cpow, u32* cinv, u32* consts, size_t n, int a, int b) {
  typedef typename F::Params P;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F g, gi;
#pragma unroll
  for (int l = 0; l < 10; l++) {
    g.l[l] = P::generator(l);
    gi.l[l] = P::generator_inv(l);
  }
  // LARGE_SUBGROUP_ROOT^(2^s 49 / n) = g^((p-1)/n): exponent by dividing p - 1 by n = 7^a 2^b limb-wise
  u32 e[10];
#pragma unroll
  for (int l = 0; l < 10; l++) e[l] = P::mod(l);
  e[0] -= 1;
  u32 d7 = a == 0 ? 1u : (a == 1 ? 7u : 49u);
  unsigned long long rem = 0;
  for (int l = 9; l >= 0; l--) {
    unsigned long long cur = (rem << 32) | e[l];
    e[l] = (u32)(cur / d7);
    rem = cur % d7;
  }
  for (int l = 0; l < 10; l++) {  // >> b
    unsigned long long v = e[l] | ((unsigned long long)(l + 1 < 10 ? e[l + 1] : 0u) << 32);
    e[l] = (u32)(v >> b);
  }
  F w = F::one();
  {
    bool started = false;
    for (int l = 9; l >= 0; l--)
      for (int bit = 31; bit >= 0; bit--) {
        if (started) w = w.sqr();
        if ((e[l] >> bit) & 1) {
          w = started ? w * g : g;
          started = true;
        }
      }
  }
  F inv7 = F::from_u32(d7).inverse();
  st_elem<F>(wpow, i, w.pow64(i));
  st_elem<F>(cpow, i, g.pow64(i));
  st_elem<F>(cinv, i, gi.pow64(i) * inv7);
  if (i == 0) {
    st_elem<F>(consts, 0, inv7);
    F gn = g.pow64(n) - F::one();
    st_elem<F>(consts, 1, gn.inverse());  // 1 / Z on the coset: the vanishing polynomial is x^n - 1 there too
  }
}

struct MixedTables {
  const u32 *wpow, *cpow, *cinv, *inv7;
};

template <class F>
__global__ void __launch_bounds__(128) mixed_split_kernel(const u32* __restrict__ src, u32* __restrict__ dst,
                                                          const u32* __restrict__ wpow, const u32* __restrict__ pre,
                                                          size_t n, size_t L, size_t stride7, int inverse) {
  // rows = n / L sub-arrays of length L = 7 M; stride7 = n / L (= 7^level): w_L = w_N^stride7
  const size_t M = L / 7;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n / 7) return;
  size_t row = gid / M, n2 = gid - row * M;
  const u32* in = src + row * L * 10;
  u32* out = dst + row * L * 10;
  F x[7];
#pragma unroll
  for (int n1 = 0; n1 < 7; n1++) {
    x[n1] = ld_elem<F>(in, M * n1 + n2);
    if (pre) x[n1] = x[n1] * ldg_elem<F>(pre, row * L + M * n1 + n2);
  }
  const size_t seventh = n / 7;
  for (int k1 = 0; k1 < 7; k1++) {
    F acc = x[0];
    for (int n1 = 1; n1 < 7; n1++) {
      size_t e = seventh * (size_t)((n1 * k1) % 7);
      if (e == 0) acc = acc + x[n1];
      else acc = acc + x[n1] * ldg_elem<F>(wpow, inverse ? n - e : e);
    }
    size_t te = (stride7 * n2 % n) * (size_t)k1 % n;
    if (te != 0) acc = acc * ldg_elem<F>(wpow, inverse ? n - te : te);
    st_elem<F>(out, (size_t)k1 * M + n2, acc);
  }
}

template <class F>
__global__ void __launch_bounds__(256) mixed_gather_kernel(const u32* __restrict__ src, u32* __restrict__ dst,
                                                           const u32* __restrict__ post, const u32* __restrict__ post_c,
                                                           size_t n, int a, int b) {
  size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  size_t row, k2;
  if (a == 1) {
    row = k % 7;
    k2 = k / 7;
  } else {
    row = (k % 7) * 7 + (k / 7) % 7;
    k2 = k / 49;
  }
  F v = ld_elem<F>(src, (row << b) + k2);
  if (post) v = v * ldg_elem<F>(post, k);
  else if (post_c) v = v * ldg_elem<F>(post_c, 0);
  st_elem<F>(dst, k, v);
}

template <class F>
static int mixed_tables_get(pcdgpu_ctx* ctx, int field, int a, int b, MixedTables* out) {
  const size_t n = (size_t)(a == 1 ? 7 : 49) << b;
  int key = 4096 + field * 1024 + a * 64 + b;
  auto it = ctx->ntt_tables.find(key);
  if (it == ctx->ntt_tables.end()) {
    NttTables t;
    void* all = nullptr;
    size_t bytes = (3 * n + 2) * 40;
    cudaError_t e = cudaMalloc(&all, bytes);
    if (e != cudaSuccess) {
      ctx->set_error("cudaMalloc(%zu) for mixed-radix NTT tables: %s", bytes, cudaGetErrorString(e));
      return PCDGPU_E_NOMEM;
    }
    t.twiddles = all;
    t.coset_pow = (char*)all + n * 40;
    t.coset_inv = (char*)all + 2 * n * 40;
    mixed_tables_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
        (u32*)t.twiddles, (u32*)t.coset_pow, (u32*)t.coset_inv, (u32*)((char*)all + 3 * n * 40), n, a, b);
    PCD_CUDA(ctx, cudaGetLastError());
    it = ctx->ntt_tables.emplace(key, t).first;
  }
  *out = MixedTables{(const u32*)it->second.twiddles, (const u32*)it->second.coset_pow,
                     (const u32*)it->second.coset_inv, (const u32*)((const char*)it->second.twiddles + 3 * n * 40)};
  return 0;
}

template <class F>
static int ntt_mixed_t(pcdgpu_ctx* ctx, int field, void* d_data, int a, int b, int inverse, int coset) {
  const size_t n = (size_t)(a == 1 ? 7 : 49) << b;
  MixedTables t;
  PCD_TRY(mixed_tables_get<F>(ctx, field, a, b, &t));
  void* tmp;
  PCD_TRY(ctx->scratch(SLOT_NTT_MIXED, n * 40, &tmp));
  int ps = ctx->prof_begin(PROF_NTT, (double)n * (8.0 * a + 0.5 * b));
  // split levels: data -> tmp (-> data)
  void* cur = d_data;
  void* other = tmp;
  size_t L = n, stride7 = 1;
  for (int lvl = 0; lvl < a; lvl++) {
    const u32* pre = (lvl == 0 && coset && !inverse) ? t.cpow : nullptr;
    mixed_split_kernel<F><<<(unsigned)((n / 7 + 127) / 128), 128, 0, ctx->stream>>>((const u32*)cur, (u32*)other, t.wpow,
                                                                                 pre, n, L, stride7, inverse);
    PCD_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    std::swap(cur, other);
    L /= 7;
    stride7 *= 7;
  }
  // 7^a rows of length 2^b: batched radix-2 transform, in place in `cur`
  ctx->prof_end(ps);
  if (b > 0) PCD_TRY((ntt_run_t<F>(ctx, field, cur, b, inverse, 0, (int)(n >> b))));
  ps = ctx->prof_begin(PROF_NTT, 0.0);
  const u32* post = (coset && inverse) ? t.cinv : nullptr;
  const u32* post_c = (inverse && !coset) ? t.inv7 : nullptr;
  mixed_gather_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const u32*)cur, (u32*)other, post, post_c,
                                                                            n, a, b);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  if (other != d_data) PCD_CUDA(ctx, cudaMemcpyAsync(d_data, other, n * 40, cudaMemcpyDeviceToDevice, ctx->stream));
  ctx->prof_end(ps);
  return 0;
}

// transform on the domain 7^a 2^b (a = 0: radix 2)
int ntt_run_general(pcdgpu_ctx* ctx, int field, void* d_data, int a, int b, int inverse, int coset) {
  if (a == 0) return ntt_run(ctx, field, d_data, b, inverse, coset);
  if (field != PCDGPU_FIELD_Q4 || a > 2 || b > 17) {
    ctx->set_error("no evaluation domain of size 7^%d 2^%d on this field", a, b);
    return PCDGPU_E_DOMAIN;
  }
  return ntt_mixed_t<FpQ4>(ctx, field, d_data, a, b, inverse, coset);
}

int ntt_domain_shape(int field, size_t min_size, size_t* n, int* a, int* b);
int ntt_zinv_general(pcdgpu_ctx* ctx, int field, size_t n, const u32** d_zinv) {
  size_t nn;
  int a, b;
  PCD_TRY(ntt_domain_shape(field, n, &nn, &a, &b));
  if (nn != n) return PCDGPU_E_DOMAIN;
  if (a == 0) {
    NttTablesDev t;
    PCD_TRY(ntt_tables(ctx, field, b, &t));
    *d_zinv = t.zinv;
    return 0;
  }
  MixedTables t;
  PCD_TRY(mixed_tables_get<FpQ4>(ctx, field, a, b, &t));
  *d_zinv = t.inv7 + 10;
  return 0;
}

// GeneralEvaluationDomain::new(min_size): radix 2 if it fits the 2-adicity, else the smallest 7^a 2^b on q4
int ntt_domain_shape(int field, size_t min_size, size_t* n, int* a, int* b) {
  int two_adicity = field == PCDGPU_FIELD_R4 ? 34 : 17;
  int lg = ilog2_ceil(min_size ? min_size : 1);
  if (lg <= two_adicity) {
    *n = (size_t)1 << lg;
    *a = 0;
    *b = lg;
    return 0;
  }
  if (field != PCDGPU_FIELD_Q4) return PCDGPU_E_DOMAIN;
  bool found = false;
  for (int aa = 0; aa <= 2; aa++)
    for (int bb = 0; bb <= 17; bb++) {
      size_t s = (size_t)(aa == 0 ? 1 : (aa == 1 ? 7 : 49)) << bb;
      if (s >= min_size && (!found || s < *n)) {
        *n = s;
        *a = aa;
        *b = bb;
        found = true;
      }
    }
  return found ? 0 : PCDGPU_E_DOMAIN;
}

// `batch` transforms of 2^log_n elements stored back to back, one launch per pass (blockIdx.y = transform)
int ntt_run_batch(pcdgpu_ctx* ctx, int field, void* d_data, int log_n, int inverse, int coset, int batch) {
  if (field == PCDGPU_FIELD_R4) return ntt_run_t<FpR4>(ctx, field, d_data, log_n, inverse, coset, batch);
  if (field == PCDGPU_FIELD_Q4) return ntt_run_t<FpQ4>(ctx, field, d_data, log_n, inverse, coset, batch);
  ctx->set_error("unknown field id %d", field);
  return PCDGPU_E_ARG;
}
int ntt_run(pcdgpu_ctx* ctx, int field, void* d_data, int log_n, int inverse, int coset) {
  if (field == PCDGPU_FIELD_R4) return ntt_run_t<FpR4>(ctx, field, d_data, log_n, inverse, coset);
  if (field == PCDGPU_FIELD_Q4) return ntt_run_t<FpQ4>(ctx, field, d_data, log_n, inverse, coset);
  ctx->set_error("unknown field id %d", field);
  return PCDGPU_E_ARG;
}
