"""The constants every transform and key depends on, pinned to the values recorded in SURVEY.md Appendix A (the
published ark-mnt4-298 / ark-mnt6-298 / libff parameter files as recalled and then verified arithmetically there):
moduli, TWO_ADIC_ROOT_OF_UNITY, LARGE_SUBGROUP_ROOT_OF_UNITY, GENERATOR, Montgomery R and R^2, INV, curve coefficients
and the MNT4 G1 generator.  A shared mistake in a root or generator would pass every GPU-vs-oracle test; it cannot pass
this one.  (The values themselves are [RECALL]+[CHK] in the survey's legend: arkworks' source is not on this box.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pcd_oracle as po  # noqa: E402

R4 = 475922286169261325753349249653048451545124878552823515553267735739164647307408490559963137
Q4 = 475922286169261325753349249653048451545124879242694725395555128576210262817955800483758081


def limbs64(v):
    return [(v >> (64 * i)) & (2 ** 64 - 1) for i in range(5)]


def test_moduli_and_montgomery_constants():
    assert po.R4 == R4 and po.Q4 == Q4
    assert limbs64(R4) == [0xbb4334a400000001, 0xfb494c07925d6ad3, 0xcaeec9635cf44194, 0xa266249da7b0548e, 0x000003bcf7bcd473]
    assert limbs64(Q4) == [0xc90cd65a71660001, 0x41a9e35e51200e12, 0xcaeec9635d1330ea, 0xa266249da7b0548e, 0x000003bcf7bcd473]
    assert limbs64((1 << 320) % R4) == [0xc3177aefffbb845c, 0x9b80c702f9961788, 0xc5df8dcdac70a85a, 0x29184098647b5197, 0x000001c1223d33c3]
    assert limbs64((1 << 640) % R4) == [0x465a743c68e0596b, 0x034f9102adb68371, 0x4bbd6dcf1e3a8386, 0x02ff00dced8e4b6d, 0x00000149bb44a342]
    assert limbs64((1 << 320) % Q4) == [0x18c31a7b5863845c, 0xe9de7a15e3b68df5, 0xc5df858728faab40, 0x29184098647b5197, 0x000001c1223d33c3]
    assert limbs64((1 << 640) % Q4) == [0x0065acec5613d220, 0xa266a1adbf2bc893, 0x66bd7673318850e1, 0x1f32e014ad38d47b, 0x00000224f0918a34]
    assert (-pow(R4, -1, 1 << 64)) % (1 << 64) == 0xbb4334a3ffffffff
    assert (-pow(Q4, -1, 1 << 64)) % (1 << 64) == 0xb071a1b67165ffff
    assert po.FR4.inv64 == 0xbb4334a3ffffffff and po.FQ4.inv64 == 0xb071a1b67165ffff


def test_roots_and_generators():
    assert po.FR4.generator == 10 and po.FQ4.generator == 17
    assert po.FR4.two_adicity == 34 and po.FQ4.two_adicity == 17
    assert po.FR4.two_adic_root == \
        120638817826913173458768829485690099845377008030891618010109772937363554409782252579816313
    assert po.FQ4.two_adic_root == \
        264706250571800080758069302369654305530125675521263976034054878017580902343339784464690243
    large = pow(17, (Q4 - 1) // ((1 << 17) * 49), Q4)
    assert large == 381811485921190977554243339163030148371175054922689353173385941180422489253833691237722982
    assert pow(large, 49, Q4) == po.FQ4.two_adic_root  # radix-2 and mixed-radix domains agree on omega
    # the oracle's mixed-radix domain uses exactly that root
    d = po.domain_mixed(po.FQ4, 2, 17)
    assert d.omega == large


def test_curves():
    assert po.MNT4_G1.a == 2 and po.MNT6_G1.a == 11
    assert po.MNT4_G1.b == 423894536526684178289416011533888240029318103673896002803341544124054745019340795360841685
    assert po.MNT6_G1.b == 106700080510851735677967319632585352256454251201367587890185989362936000262606668469523074
    gx = 60760244141852568949126569781626075788424196370144486719385562369396875346601926534016838
    gy = 363732850702582978263902770815145784459747722357071843971107674179038674942891694705904306
    assert po.generator(po.MNT4_G1) == (gx, gy)
    # twists: a' = a * nr (Fq2: (34, 0)), b' = (0, 17 b); Fq3: a' = (0, 0, 11), b' = (5 b, 0, 0)
    assert po.MNT4_G2.a == (34, 0) and po.MNT4_G2.b == (0, 17 * po.MNT4_G1.b % Q4)
    assert po.MNT6_G2.a == (0, 0, 11) and po.MNT6_G2.b == (5 * po.MNT6_G1.b % R4, 0, 0)
