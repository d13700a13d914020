"""Host-side mirror of ark-marlin's Fiat-Shamir transform for the reference's Marlin configuration
(/root/reference/tests/mnt4_marlin.rs:53-54: `FS4 = FiatShamirAlgebraicSpongeRng<Fr, Fq, PoseidonSponge<Fq>>`, `FS6` the
same with the fields swapped): ark-marlin src/fiat_shamir/mod.rs and src/fiat_shamir/poseidon/mod.rs (branch
`constraints`, un-vendored: restated from the published structure, see DESIGN.md "Marlin").

This is host logic above the C ABI -- a few hundred field multiplications per proof on challenge-sized data -- exactly the
part that stays in Rust when the reference is patched to call libpcdgpu.so; it is written here with Python integers.
Names follow ark-marlin: `PoseidonSponge::{absorb, squeeze}`, `FiatShamirAlgebraicSpongeRng::{absorb_bytes,
absorb_native_field_elements, absorb_nonnative_field_elements, squeeze_nonnative_field_elements,
squeeze_128_bits_nonnative_field_elements}`."""
from __future__ import annotations

from typing import Iterable, List

from .synthetic import FIELD_P

MODULUS_BITS = 298
_R_INV = {f: pow(1 << 320, -1, p) for f, p in FIELD_P.items()}


class ChaChaRng:
    """rand_chacha::ChaChaRng (ChaCha20 keystream words, block counter from 0) behind `SeedableRng::seed_from_u64`"""

    _SIGMA = (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)

    def __init__(self, seed_u64: int):
        words, s = [], seed_u64 & (2 ** 64 - 1)
        for _ in range(8):  # rand_core: PCG32 fills the 32-byte seed four bytes at a time
            s = (s * 6364136223846793005 + 11634580027462260723) % 2 ** 64
            x = (((s >> 18) ^ s) >> 27) % 2 ** 32
            rot = s >> 59
            words.append(((x >> rot) | (x << (-rot % 32))) % 2 ** 32)
        self._key = words
        self._block_no = 0
        self._words: List[int] = []

    @staticmethod
    def _rotl(v, k):
        return ((v << k) | (v >> (32 - k))) & 0xFFFFFFFF

    def _refill(self):
        init = list(self._SIGMA) + self._key + [self._block_no & 0xFFFFFFFF, self._block_no >> 32, 0, 0]
        w = list(init)
        rotl = self._rotl
        for _ in range(10):
            for a, b, c, d in ((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
                               (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14)):
                w[a] = (w[a] + w[b]) & 0xFFFFFFFF
                w[d] = rotl(w[d] ^ w[a], 16)
                w[c] = (w[c] + w[d]) & 0xFFFFFFFF
                w[b] = rotl(w[b] ^ w[c], 12)
                w[a] = (w[a] + w[b]) & 0xFFFFFFFF
                w[d] = rotl(w[d] ^ w[a], 8)
                w[c] = (w[c] + w[d]) & 0xFFFFFFFF
                w[b] = rotl(w[b] ^ w[c], 7)
        self._words = [(x + y) & 0xFFFFFFFF for x, y in zip(w, init)]
        self._block_no += 1

    def next_u64(self) -> int:
        out = 0
        for half in range(2):
            if not self._words:
                self._refill()
            out |= self._words.pop(0) << (32 * half)
        return out

    def field_element(self, field: int) -> int:
        """ark-ff `Fp320::rand`: 5 x u64, top limb masked to 298 bits, read as the Montgomery representation"""
        p = FIELD_P[field]
        while True:
            raw = sum(self.next_u64() << (64 * i) for i in range(5)) & ((1 << MODULUS_BITS) - 1)
            if raw < p:
                return raw * _R_INV[field] % p


class PoseidonSponge:
    """ark-marlin fiat_shamir::poseidon::PoseidonSponge<CF>: width 3 (rate 2 + capacity 1), 8 full + 31 partial rounds,
    x^17, the 0/1 circulant MDS, round constants from ChaChaRng::seed_from_u64(123456789)."""

    RATE, WIDTH, FULL, PARTIAL, ALPHA = 2, 3, 8, 31, 17
    MDS = ((1, 0, 1), (1, 1, 0), (0, 1, 1))
    _constants = {}

    def __init__(self, field: int):
        self.field, self.p = field, FIELD_P[field]
        if field not in self._constants:
            rng = ChaChaRng(123456789)
            self._constants[field] = [[rng.field_element(field) for _ in range(self.WIDTH)]
                                      for _ in range(self.FULL + self.PARTIAL)]
        self.ark = self._constants[field]
        self.state = [0] * self.WIDTH
        self.absorbing, self.pos = True, 0

    def _permute(self):
        p, s = self.p, self.state
        first_partial, first_tail = self.FULL // 2, self.FULL // 2 + self.PARTIAL
        for r, consts in enumerate(self.ark):
            s = [(v + c) % p for v, c in zip(s, consts)]
            if first_partial <= r < first_tail:
                s[0] = pow(s[0], self.ALPHA, p)
            else:
                s = [pow(v, self.ALPHA, p) for v in s]
            s = [sum(m * v for m, v in zip(row, s)) % p for row in self.MDS]
        self.state = s

    def absorb(self, elems: Iterable[int]):
        elems = [e % self.p for e in elems]
        if not elems:
            return
        if not self.absorbing:  # squeezing -> absorbing: permute, start at the first rate element
            self._permute()
            self.absorbing, self.pos = True, 0
        for e in elems:
            if self.pos == self.RATE:
                self._permute()
                self.pos = 0
            self.state[self.pos] = (self.state[self.pos] + e) % self.p
            self.pos += 1

    def squeeze(self, n: int) -> List[int]:
        out: List[int] = []
        if n == 0:
            return out
        if self.absorbing:
            self._permute()
            self.absorbing, self.pos = False, 0
        while len(out) < n:
            if self.pos == self.RATE:
                self._permute()
                self.pos = 0
            out.append(self.state[self.pos])
            self.pos += 1
        return out


#: limb shape of a non-native element on its way into the sponge.  ark-nonnative-field derives it by a cost search
#: (`get_params(298, 298, OptimizationType::Weight)`) that is not restated; 10 x 30 bits is this mirror's stand-in
NONNATIVE_LIMBS, NONNATIVE_LIMB_BITS = 10, 30


class FiatShamirAlgebraicSpongeRng:
    def __init__(self, field: int, sponge_field: int):
        self.field, self.sponge_field = field, sponge_field
        self.sponge = PoseidonSponge(sponge_field)

    def absorb_native_field_elements(self, elems: Iterable[int]):
        self.sponge.absorb(elems)

    def absorb_nonnative_field_elements(self, elems: Iterable[int]):
        """limbs most significant first; normal-form limbs are bounded by `limb bits + 1`, neighbours are packed in pairs
        (compress_elements) -- they always fit below the sponge field's capacity here"""
        width = NONNATIVE_LIMB_BITS + 1
        assert 2 * width <= MODULUS_BITS - 1
        flat: List[int] = []
        for e in elems:
            e %= FIELD_P[self.field]
            flat += [(e >> (NONNATIVE_LIMB_BITS * k)) & ((1 << NONNATIVE_LIMB_BITS) - 1)
                     for k in range(NONNATIVE_LIMBS - 1, -1, -1)]
        packed = [(flat[i] << width) + flat[i + 1] if i + 1 < len(flat) else flat[i] for i in range(0, len(flat), 2)]
        self.sponge.absorb(packed)

    def absorb_bytes(self, data: bytes):
        chunk = MODULUS_BITS - 128
        bits = "".join(format(b, "08b")[::-1] for b in data)  # each byte least significant bit first
        self.sponge.absorb([int(bits[i:i + chunk], 2) for i in range(0, len(bits), chunk)])

    def _squeeze_bits(self, count: int) -> str:
        per = MODULUS_BITS - 1
        need = -(-count // per)
        return "".join(format(e, "0%db" % MODULUS_BITS)[-per:] for e in self.sponge.squeeze(need))[:count]

    def _squeeze_elements(self, n: int, bits_each: int) -> List[int]:
        stream = self._squeeze_bits(n * bits_each)
        p = FIELD_P[self.field]
        return [int(stream[k * bits_each:(k + 1) * bits_each][::-1], 2) % p for k in range(n)]  # first bit = 2^0

    def squeeze_nonnative_field_elements(self, n: int) -> List[int]:
        return self._squeeze_elements(n, MODULUS_BITS - 1)

    def squeeze_128_bits_nonnative_field_elements(self, n: int) -> List[int]:
        return self._squeeze_elements(n, 128)

    def squeeze_native_field_elements(self, n: int) -> List[int]:
        return self.sponge.squeeze(n)
