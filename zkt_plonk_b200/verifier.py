"""Host-side verifier through the C ABI (csrc/verify.cu): `Proof::verify` of the reference
(plonk-core/src/proof_system/proof.rs:285-503) with PC::check as a product of pairings on the selected curve (BN254 unless
`field.use_curve` chose BLS12-381 / BLS12-377: each curve's build of the library has its own pairing).  No GPU is used."""
import ctypes

import numpy as np

from . import _lib, field
from .prover import VerifierKey, ints_to_mont_array
from .transcript import TRANSCRIPTS


def _lib_now():
    return _lib.lib(field.CURVE)


def _fq_mont_limbs(v):
    m = int(v) * field.MONT_RQ % field.Q_MOD
    return [(m >> (64 * k)) & (2**64 - 1) for k in range(field.FQ_WORDS)]


def g1_array(pt):
    """(x, y) canonical ints or None -> (2 * FQ_WORDS,) uint64 Montgomery affine, identity = zeros."""
    w = field.FQ_WORDS
    out = np.zeros(2 * w, dtype=np.uint64)
    if pt is not None:
        out[:w], out[w:] = _fq_mont_limbs(pt[0]), _fq_mont_limbs(pt[1])
    return out


def g2_array(pt):
    """((x0, x1), (y0, y1)) canonical ints -> (4 * FQ_WORDS,) uint64: x.c0 x.c1 y.c0 y.c1 in Montgomery form (arkworks' G2Affine)."""
    w = field.FQ_WORDS
    if isinstance(pt, np.ndarray):                       # already in the ABI's form (e.g. from keyfile.cvk_read)
        return np.ascontiguousarray(pt, dtype=np.uint64).reshape(4 * w)
    out = np.zeros(4 * w, dtype=np.uint64)
    if pt is not None:
        (x0, x1), (y0, y1) = pt
        for k, v in enumerate((x0, x1, y0, y1)):
            out[w * k: w * k + w] = _fq_mont_limbs(v)
    return out


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# alt_bn128 G2 generator (EIP-197), ((x.c0, x.c1), (y.c0, y.c1))
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))
# a point of order r on the twist per curve, for synthetic SRSs (KZG10::setup draws h at random): the standard G2 generator
# of BLS12-381; on BLS12-377 the point x = 2 + i with its cofactor cleared
G2_POINTS = {
    "bn254": G2_GEN,
    "bls12_381": ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
                   0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
                  (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
                   0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
    "bls12_377": ((0x6f72205595a839df693176b247c2fa251f7e02a29061e50540dc9e1c2bf1957bf1bab2288c257c2cb36b58f2418bc9,
                   0x138c24b2b4e17888beed0a9802aac837cdea39890effe00072f754ecb0152dd6cb524f281298966dbaeca23d3e462b8),
                  (0x16235fdea6c3faf2a83d3730f6ab2c033ef6c2739002946f7dc48e4688bca1af1c9b417d58220817e0dc644b5e7d916,
                   0x707ac6cc7d192827fc54eb83267f3bed8511bd3c74f63a1ea75eabb66476769c8786f2af2a75166f33142379b4963c)),
}


def g2_mul(k, pt=None):
    """k * pt on G2 through the C ABI (zkb_g2_mul); returns the Montgomery array the verifier takes."""
    pt = G2_POINTS[field.CURVE] if pt is None else pt
    k = int(k)
    sc = np.array([(k >> (64 * j)) & (2**64 - 1) for j in range(4)], dtype=np.uint64)
    out = np.zeros(4 * field.FQ_WORDS, dtype=np.uint64)
    rc = _lib_now().zkb_g2_mul(_vp(g2_array(pt)), _vp(sc), _vp(out))
    if rc != 0:
        raise _lib.ZkbError(rc, "zkb_g2_mul: point not on the twist")
    return out


def make_cvk(tau):
    """(h, beta_h) = (H, tau * H) for a synthetic SRS with known trapdoor (kzg10::setup draws h at random; any h gives the
    same accept / reject decisions).  Product code: bench.py verifies every timed proof with it."""
    return g2_array(G2_POINTS[field.CURVE]), g2_mul(tau)


def verify(vk, proof_bytes, pub_inputs, cvk_g2, transcript="merlin"):
    """vk: prover.VerifierKey; proof_bytes: the serialised proof (802 bytes on BN254, 1010 on the BLS12 curves); pub_inputs: canonical ints, one per vk.pi_roots entry;
    cvk_g2 = (h, beta_h) as Fq2 coordinate tuples of canonical ints or as the (16,) arrays keyfile.cvk_read returns.  Returns 0 (accepted), 1 or 2 (failing step); raises on malformed input."""
    assert len(proof_bytes) == 13 * field.FQ_BYTES + 2 + 12 * 32 and len(pub_inputs) == len(vk.pi_roots), "invalid length of public inputs"
    xy = np.stack([g1_array(vk.commits[name]) for name in VerifierKey.ORDER])
    inf = (ctypes.c_int * 10)(*[int(vk.commits[name] is None) for name in VerifierKey.ORDER])
    roots = ints_to_mont_array(vk.pi_roots) if vk.pi_roots else np.zeros((1, 4), dtype=np.uint64)
    pub = ints_to_mont_array(list(pub_inputs)) if pub_inputs else np.zeros((1, 4), dtype=np.uint64)
    raw = np.frombuffer(bytes(proof_bytes), dtype=np.uint8).copy()
    h, bh = g2_array(cvk_g2[0]), g2_array(cvk_g2[1])
    rc = _lib_now().zkb_plonk_verify(vk.n, _vp(roots), len(vk.pi_roots), _vp(xy), inf, _vp(pub), _vp(raw), _vp(h), _vp(bh),
                                     TRANSCRIPTS[transcript][0])
    if rc < 0:
        raise _lib.ZkbError(rc, "zkb_plonk_verify: malformed verifier key, proof or G2 elements")
    return rc


def pairing(p1, q2):
    """e(P, Q): 12 canonical Fq12 coefficients in the w-basis (ints)."""
    w = field.FQ_WORDS
    out = np.zeros((12, w), dtype=np.uint64)
    rc = _lib_now().zkb_pairing(_vp(g1_array(p1)), _vp(g2_array(q2)), _vp(out))
    if rc != 0:
        raise _lib.ZkbError(rc, "zkb_pairing: point not on its curve")
    return [sum(int(out[k, j]) << (64 * j) for j in range(w)) for k in range(12)]


def pairing_product_is_one(pairs):
    g1 = np.stack([g1_array(p) for p, _ in pairs])
    g2 = np.stack([g2_array(q) for _, q in pairs])
    one = ctypes.c_int(0)
    rc = _lib_now().zkb_pairing_product_is_one(_vp(g1), _vp(g2), len(pairs), ctypes.byref(one))
    if rc != 0:
        raise _lib.ZkbError(rc, "zkb_pairing_product_is_one: point not on its curve")
    return bool(one.value)
