// groth16.cuh -- device-resident objects behind the opaque ABI handles, and the internal
// interface of groth16.cu.
#pragma once
#include "common.cuh"
#include "vecio.cuh"

struct CsrDev {
  const u32* row_ptr;  // m + 1
  const u32* col;      // nnz
  const u32* val;      // nnz x 10 words, Montgomery
};

struct pcdgpu_r1cs {
  pcdgpu_ctx* ctx;
  int pairing;
  size_t m, num_inputs, num_witness, n;  // n = domain size = 7^dom_a 2^dom_b (GeneralEvaluationDomain::new)
  int dom_a, dom_b;
  CsrDev A, B, C;
  void* storage;  // one allocation holding all nine arrays
  unsigned long long uid;  // unique per upload (keys of the context's proof graphs; an address can be reused)
};

struct pcdgpu_bases {
  pcdgpu_ctx* ctx;
  int curve;
  size_t n;       // points
  void* points;   // n affine points (device)
  int c, nwin;    // precomputed table parameters (c = 0: none)
  void* table;    // nwin x n affine points: table[j * n + i] = 2^(c j) * points[i]
};

struct pcdgpu_pk {
  pcdgpu_ctx* ctx;
  int pairing;
  size_t num_vars, num_inputs, h_len;
  // query vectors with their constant points appended (see pcdgpu_pk_upload)
  pcdgpu_bases *a_query, *b_g1_query, *b_g2_query, *h_query, *l_query;
  // sharded keys (pcdgpu_pk_upload_sharded): this rank holds points [lo, hi) of each query; the constant points ride
  // with rank 0.  world 1: the whole key.
  int shard_rank, shard_world;
  size_t v_lo, v_hi;  // of the num_vars - 1 ordinary points of a / b_g1 / b_g2
  size_t h_lo, h_hi;  // of h_query
  size_t l_lo, l_hi;  // of l_query
  unsigned long long uid;  // unique per upload (see pcdgpu_r1cs)
};

struct pcdgpu_gm17_pk {
  pcdgpu_ctx* ctx;
  int pairing;
  size_t num_sap_vars, num_inputs, h_len;
  // query vectors with their constant points appended (see pcdgpu_gm17_pk_upload)
  pcdgpu_bases *a_query, *b_query, *c_query_1, *c_query_2, *g_gamma2_z_t;
};

#if defined(__CUDACC__)
template <class F>
__device__ __forceinline__ F ld10(const u32* g, size_t idx) {
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
__device__ __forceinline__ void st10(u32* g, size_t idx, const F& a) {
  uint2* p = reinterpret_cast<uint2*>(g + idx * 10);
#pragma unroll
  for (int i = 0; i < 5; i++) p[i] = make_uint2(a.l[2 * i], a.l[2 * i + 1]);
}

#endif

// shared by api.cu and gm17.cu
int pcd_g1_of(int pairing);
int pcd_g2_of(int pairing);
// MSM over points [offset, offset + n) of a resident vector followed by its last n_extra points (the
// per-proof constant pairs appended at key upload) with plain scalars d_extra; result: one xyzz point
int bases_msm(pcdgpu_ctx* ctx, const pcdgpu_bases* b, size_t offset, const void* d_scalars, int scalars_mont, size_t n,
              const void* d_extra, size_t n_extra, void* d_out_xyzz);

// a, b, c <- matrices x z; h = coset_ifft((coset_fft(ifft a) * coset_fft(ifft b) - coset_fft(ifft c)) / Z).
// *d_h points into the context's scratch (n elements, Montgomery form).
int witness_map_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, const void* d_z, void** d_h);
int qap_vector_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, int which, const void* d_z, void* d_out);
int qap_combine_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, void* d_a, const void* d_b, const void* d_c);
// extras <- {r, 1, 1, s, 1, 1, -(r s) mod p, rs, s, s, rs, r, r} as plain integers (the scalars of the constant pairs)
int groth16_prepare(pcdgpu_ctx* ctx, int pairing, const u32* d_rs, u32* d_extras);
// the proof's tail, on the context's current lane (ctx->cur()): sums1 = {h_acc, l_acc - (r s) delta, T | s g_a, r g1_b, g_a, g1_b}
int groth16_sum_partials(pcdgpu_ctx* ctx, int pairing, const void* p1, const void* p2, int world, int n1, int n2,
                         void* out1, void* out2);
int groth16_straus(pcdgpu_ctx* ctx, int pairing, const u32* d_rs, void* sums1);      // T = s g_a + r g1_b
int point_to_affine(pcdgpu_ctx* ctx, int curve, const void* src_xyzz, size_t idx, void* dst_affine);
int groth16_finish(pcdgpu_ctx* ctx, int pairing, const void* sums1, void* d_out_c, int nterms = 3);  // C = h + l' + T (+ ...), affine
int groth16_scale(pcdgpu_ctx* ctx, int pairing, const void* d_z, size_t n, const u32* d_rs, void* d_sz, void* d_rz);
int groth16_serialize(pcdgpu_ctx* ctx, int pairing, const void* d_proof, unsigned char* d_out);
int bench_imad(pcdgpu_ctx* ctx, int modmul, int iters, double* ops_per_s, double* ms_out);
// comm.cu: all-gather `bytes` bytes per rank (rank-major result) on stream st; group brackets (ncclGroupStart / End)
int comm_allgather(pcdgpu_ctx* ctx, const void* d_send, void* d_recv, size_t bytes, cudaStream_t st);
int comm_group(pcdgpu_ctx* ctx, bool start);
