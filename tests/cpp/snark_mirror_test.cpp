// snark_mirror_test.cpp -- drives include/pcdgpu_snark.hpp the way the reference's tests drive
// `Groth16::<E>::prove` (tests/mnt4_groth16.rs:23-30,86): index the key, prove with an rng whose draws
// are r then s, serialize.  Input: a binary fixture written by tests/test_cpp_mirror.py from
// tests/golden/groth16.json; output: the proof's affine limbs and canonical bytes.
// Exit codes: 0 ok, 2 backend unavailable (no GPU: there is no CPU fallback), 1 anything else.
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "pcdgpu_snark.hpp"

using namespace pcdgpu;

extern "C" int cudaGetDeviceCount(int* count);  // libcudart (the test links it; no CUDA headers needed for this)

static std::vector<uint64_t> g_data;
static size_t g_pos = 0;
static uint64_t rd() { return g_data[g_pos++]; }
template <size_t N> static std::array<uint64_t, N> rda() {
  std::array<uint64_t, N> a;
  for (size_t i = 0; i < N; i++) a[i] = rd();
  return a;
}

struct FixtureRng {  // the caller's rng: yields the fixture's r, then s (GM17: d1, d2, r)
  Fr draws[3];
  int next = 0;
  Fr next_scalar(int) { return draws[next++]; }
};

template <class E>
static int run(FILE* out) {
  SynthesizedCircuit circ;
  ConstraintMatrices& m = circ.matrices;
  m.num_constraints = rd();
  m.num_instance_variables = rd();
  m.num_witness_variables = rd();
  std::vector<std::vector<std::pair<Fr, size_t>>>* mats[3] = {&m.a, &m.b, &m.c};
  for (int k = 0; k < 3; k++) {
    mats[k]->resize(m.num_constraints);
    for (size_t i = 0; i < m.num_constraints; i++) {
      size_t cnt = rd();
      for (size_t e = 0; e < cnt; e++) {
        Fr co = rda<5>();
        size_t col = rd();
        (*mats[k])[i].push_back({co, col});
      }
    }
  }
  for (size_t i = 0; i < m.num_instance_variables; i++) circ.instance_assignment.push_back(rda<5>());
  for (size_t i = 0; i < m.num_witness_variables; i++) circ.witness_assignment.push_back(rda<5>());
  FixtureRng rng;
  rng.draws[0] = rda<5>();
  rng.draws[1] = rda<5>();
  ProvingKey<E> pk;
  pk.vk.alpha_g1 = rda<E::G1_LIMBS>();
  pk.beta_g1 = rda<E::G1_LIMBS>();
  pk.delta_g1 = rda<E::G1_LIMBS>();
  pk.vk.beta_g2 = rda<E::G2_LIMBS>();
  pk.vk.delta_g2 = rda<E::G2_LIMBS>();
  size_t nv = m.num_instance_variables + m.num_witness_variables;
  for (size_t i = 0; i < nv; i++) pk.a_query.push_back(rda<E::G1_LIMBS>());
  for (size_t i = 0; i < nv; i++) pk.b_g1_query.push_back(rda<E::G1_LIMBS>());
  for (size_t i = 0; i < nv; i++) pk.b_g2_query.push_back(rda<E::G2_LIMBS>());
  size_t hl = rd();
  for (size_t i = 0; i < hl; i++) pk.h_query.push_back(rda<E::G1_LIMBS>());
  for (size_t i = 0; i < m.num_witness_variables; i++) pk.l_query.push_back(rda<E::G1_LIMBS>());

  // the parts of the trait that stay on the CPU report Unsupported instead of pretending
  auto setup = Groth16<E>::circuit_specific_setup(circ, rng);
  if (setup.is_ok() || setup.error.kind != ErrorKind::Unsupported) return 1;

  typename Groth16<E>::Index idx;
  auto ok = Groth16<E>::index(pk, m, &idx, true);
  if (!ok) {
    fprintf(stderr, "index: %s (code %d)\n", ok.error.message.c_str(), ok.error.code);
    return ok.error.code == PCDGPU_E_NODEVICE ? 2 : 1;
  }
  auto proof = Groth16<E>::prove(idx, circ, rng);
  if (!proof) {
    fprintf(stderr, "prove: %s\n", proof.error.message.c_str());
    return 1;
  }
  auto bytes = Groth16<E>::serialize(proof.value);
  if (!bytes || bytes.value.size() != E::PROOF_BYTES) return 1;
  fwrite(proof.value.a.data(), 8, E::G1_LIMBS, out);
  fwrite(proof.value.b.data(), 8, E::G2_LIMBS, out);
  fwrite(proof.value.c.data(), 8, E::G1_LIMBS, out);
  fwrite(bytes.value.data(), 1, bytes.value.size(), out);
  // ONE proof over the GPUs of the box through the sharded entry points (pcdgpu_comm_init, pcdgpu_pk_upload_sharded,
  // pcdgpu_groth16_prove_sharded), one host thread per GPU as a Rust caller would run them: world = min(GPUs, 2); on a
  // one-GPU box the same calls with world 1.  Every rank must return the single-GPU proof bit for bit.
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != 0 || ndev < 1) return 1;
    const int world = ndev >= 2 ? 2 : 1;
    unsigned char id[PCDGPU_COMM_ID_BYTES] = {0};
    if (world > 1 && pcdgpu_comm_unique_id(id) != PCDGPU_OK) return 1;
    std::vector<int> rcs(world, 1);
    std::vector<std::thread> ts;
    for (int rank = 0; rank < world; rank++)
      ts.emplace_back([&, rank]() {
        typename Groth16<E>::Index sidx;
        auto up = Groth16<E>::index_sharded(pk, m, &sidx, rank, world > 1 ? id : nullptr, rank, world, true);
        if (!up) {
          fprintf(stderr, "index_sharded (rank %d of %d): %s\n", rank, world, up.error.message.c_str());
          return;
        }
        auto sp = Groth16<E>::create_proof_sharded(sidx, circ, rng.draws[0], rng.draws[1], rank);
        if (!sp) {
          fprintf(stderr, "create_proof_sharded (rank %d of %d): %s\n", rank, world, sp.error.message.c_str());
          return;
        }
        rcs[rank] = (sp.value.a == proof.value.a && sp.value.b == proof.value.b && sp.value.c == proof.value.c) ? 0 : 3;
      });
    for (auto& t : ts) t.join();
    for (int rc : rcs)
      if (rc != 0) {
        fprintf(stderr, "sharded proof over %d GPU(s): rank returned %d\n", world, rc);
        return 1;
      }
  }
  // a wrong-length assignment is an error, not a crash
  circ.witness_assignment.pop_back();
  auto bad = Groth16<E>::create_proof_with_reduction(idx, circ, rng.draws[0], rng.draws[1]);
  if (bad.is_ok() || bad.error.kind != ErrorKind::AssignmentMissing) return 1;
  return 0;
}

// GM17 fixture (tests/golden/gm17.json): `GM17::<E>::prove` as tests/mnt4_gm17.rs:27-28 binds it
template <class E>
static int run_gm17(FILE* out) {
  SynthesizedCircuit circ;
  ConstraintMatrices& m = circ.matrices;
  m.num_constraints = rd();
  m.num_instance_variables = rd();
  m.num_witness_variables = rd();
  std::vector<std::vector<std::pair<Fr, size_t>>>* mats[3] = {&m.a, &m.b, &m.c};
  for (int k = 0; k < 3; k++) {
    mats[k]->resize(m.num_constraints);
    for (size_t i = 0; i < m.num_constraints; i++) {
      size_t cnt = rd();
      for (size_t e = 0; e < cnt; e++) {
        Fr co = rda<5>();
        size_t col = rd();
        (*mats[k])[i].push_back({co, col});
      }
    }
  }
  for (size_t i = 0; i < m.num_instance_variables; i++) circ.instance_assignment.push_back(rda<5>());
  for (size_t i = 0; i < m.num_witness_variables; i++) circ.witness_assignment.push_back(rda<5>());
  FixtureRng rng;
  for (int i = 0; i < 3; i++) rng.draws[i] = rda<5>();
  GM17ProvingKey<E> pk;
  size_t nsap = rd(), hlen = rd();
  for (size_t i = 0; i < nsap; i++) pk.a_query.push_back(rda<E::G1_LIMBS>());
  for (size_t i = 0; i < nsap; i++) pk.b_query.push_back(rda<E::G2_LIMBS>());
  for (size_t i = 0; i < nsap - m.num_instance_variables; i++) pk.c_query_1.push_back(rda<E::G1_LIMBS>());
  for (size_t i = 0; i < nsap; i++) pk.c_query_2.push_back(rda<E::G1_LIMBS>());
  for (size_t i = 0; i < hlen; i++) pk.g_gamma2_z_t.push_back(rda<E::G1_LIMBS>());
  pk.g_gamma_z = rda<E::G1_LIMBS>();
  pk.h_gamma_z = rda<E::G2_LIMBS>();
  pk.g_ab_gamma_z = rda<E::G1_LIMBS>();
  pk.g_gamma2_z2 = rda<E::G1_LIMBS>();
  auto setup = GM17<E>::circuit_specific_setup(circ, rng);
  if (setup.is_ok() || setup.error.kind != ErrorKind::Unsupported) return 1;
  typename GM17<E>::Index idx;
  auto ok = GM17<E>::index(pk, m, &idx, true);
  if (!ok) {
    fprintf(stderr, "gm17 index: %s (code %d)\n", ok.error.message.c_str(), ok.error.code);
    return ok.error.code == PCDGPU_E_NODEVICE ? 2 : 1;
  }
  auto proof = GM17<E>::prove(idx, circ, rng);
  if (!proof) {
    fprintf(stderr, "gm17 prove: %s\n", proof.error.message.c_str());
    return 1;
  }
  auto bytes = GM17<E>::serialize(proof.value);
  if (!bytes || bytes.value.size() != E::PROOF_BYTES) return 1;
  fwrite(proof.value.a.data(), 8, E::G1_LIMBS, out);
  fwrite(proof.value.b.data(), 8, E::G2_LIMBS, out);
  fwrite(proof.value.c.data(), 8, E::G1_LIMBS, out);
  fwrite(bytes.value.data(), 1, bytes.value.size(), out);
  // a key whose lengths do not match the SAP is an error, not a crash
  pk.c_query_1.pop_back();
  typename GM17<E>::Index bad_idx;
  auto bad = GM17<E>::index(pk, m, &bad_idx, false);
  if (bad.is_ok() || bad.error.kind != ErrorKind::MalformedKey) return 1;
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  g_data.resize(n / 8);
  if (fread(g_data.data(), 8, g_data.size(), f) != g_data.size()) return 1;
  fclose(f);
  FILE* out = fopen(argv[2], "wb");
  if (!out) return 1;
  uint64_t head = rd();  // pairing | scheme << 8 (0 = Groth16, 1 = GM17)
  uint64_t pairing = head & 0xff, scheme = head >> 8;
  int rc;
  if (scheme == 0) rc = pairing == 0 ? run<MNT4_298>(out) : run<MNT6_298>(out);
  else rc = pairing == 0 ? run_gm17<MNT4_298>(out) : run_gm17<MNT6_298>(out);
  fclose(out);
  return rc;
}
