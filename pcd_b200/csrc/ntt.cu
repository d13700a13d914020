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
// (bank-conflict free for the butterflies) and runs the r stages there.  The first pass folds the
// bit reversal into its (strided) load and writes contiguous rows; later passes work in place.
// Coset scaling (g^i before a forward transform, g^-i / n after an inverse one) is fused into
// the first load / last store.  HBM traffic per pass = one read + one write of the vector.
#include "common.cuh"
#include "ntt.cuh"

static constexpr int NTT_THREADS = 256;
static constexpr int NTT_TILE_LOG = 11;  // elements per CTA tile (2^11 x 40 B = 80 KiB)
static constexpr int NTT_MAX_R = 8;

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
};

template <class F>
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(NttPass a) {
  extern __shared__ u32 sm[];
  const int W = 1 << a.logW;
  const int R = 1 << a.r;
  const int pitch = W + 1;            // row pitch (odd for W >= 2: column reads are conflict free)
  const int plane = R * pitch;        // words per limb plane
  const int tile = 1 << (a.r + a.logW);
  const int k = a.log_n;
  const size_t b = blockIdx.x;
  const int tid = threadIdx.x;

  size_t lowbase = 0, high = 0;
  if (!a.first) {
    int lb = a.s - a.logW;  // tiles per "low" range
    lowbase = (b & (((size_t)1 << lb) - 1)) << a.logW;
    high = b >> lb;
  }
  // ---- load ------------------------------------------------------------------------------
  for (int e = tid; e < tile; e += NTT_THREADS) {
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
    u32* dst = sm + j * pitch + c;
#pragma unroll
    for (int l = 0; l < 10; l++) dst[l * plane] = v.l[l];
  }
  // ---- r butterfly stages in shared memory -------------------------------------------------
  const size_t half_n = (size_t)1 << (k - 1);
  for (int u = 0; u < a.r; u++) {
    __syncthreads();
    const int sh = k - a.s - u - 1;
    for (int q = tid; q < (tile >> 1); q += NTT_THREADS) {
      int c = q & (W - 1), jj = q >> a.logW;
      int lo = jj & ((1 << u) - 1);
      int j = ((jj >> u) << (u + 1)) | lo;
      u32* p1 = sm + j * pitch + c;
      u32* p2 = p1 + (pitch << u);
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
  for (int e = tid; e < tile; e += NTT_THREADS) {
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
    const u32* srcp = sm + j * pitch + c;
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
static int ntt_run_t(pcdgpu_ctx* ctx, int field, void* d_data, int log_n, int inverse, int coset) {
  if (log_n > F::Params::TWO_ADICITY) {
    ctx->set_error("radix-2 domain 2^%d exceeds the field's 2-adicity %d", log_n, F::Params::TWO_ADICITY);
    return PCDGPU_E_DOMAIN;
  }
  if (log_n == 0) return 0;  // n = 1: every flavour is the identity
  NttTablesDev t;
  PCD_TRY(ntt_tables(ctx, field, log_n, &t));
  const int max_smem = 10 * (1 << NTT_MAX_R) * ((1 << (NTT_TILE_LOG - NTT_MAX_R)) + 1) * 4;  // r = 8, W = 8
  PCD_CUDA(ctx, cudaFuncSetAttribute(ntt_pass_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  int passes = log_n <= 10 ? 1 : (log_n + NTT_MAX_R - 1) / NTT_MAX_R;
  void* scratch = nullptr;
  if (passes > 1) PCD_TRY(ctx->scratch(SLOT_NTT, ((size_t)40) << log_n, &scratch));
  int s = 0;
  int ps = ctx->prof_begin(PROF_NTT, (double)log_n * (double)((size_t)1 << (log_n - 1)));
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
    int R = 1 << r, W = 1 << a.logW;
    size_t smem = (size_t)10 * R * (W + 1) * 4;
    unsigned grid = (unsigned)(((size_t)1 << log_n) >> (r + a.logW));
    ntt_pass_kernel<F><<<grid, NTT_THREADS, smem, ctx->stream>>>(a);
    PCD_CUDA(ctx, cudaGetLastError());
    s += r;
  }
  ctx->prof_end(ps);
  return 0;
}

int ntt_run(pcdgpu_ctx* ctx, int field, void* d_data, int log_n, int inverse, int coset) {
  if (field == PCDGPU_FIELD_R4) return ntt_run_t<FpR4>(ctx, field, d_data, log_n, inverse, coset);
  if (field == PCDGPU_FIELD_Q4) return ntt_run_t<FpQ4>(ctx, field, d_data, log_n, inverse, coset);
  ctx->set_error("unknown field id %d", field);
  return PCDGPU_E_ARG;
}
