"""Host-side verifier through the C ABI (csrc/verify.cu): `Proof::verify` of the reference
(plonk-core/src/proof_system/proof.rs:285-503) with PC::check as a product of BN254 pairings.  No GPU is used."""
import ctypes

import numpy as np

from . import _lib
from .prover import VerifierKey, ints_to_mont_array
from .transcript import TRANSCRIPTS

Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def _fq_mont_limbs(v):
    m = (int(v) << 256) % Q_MOD
    return [(m >> (64 * k)) & (2**64 - 1) for k in range(4)]


def g1_array(pt):
    """(x, y) canonical ints or None -> (8,) uint64 Montgomery affine, identity = zeros."""
    out = np.zeros(8, dtype=np.uint64)
    if pt is not None:
        out[:4], out[4:] = _fq_mont_limbs(pt[0]), _fq_mont_limbs(pt[1])
    return out


def g2_array(pt):
    """((x0, x1), (y0, y1)) canonical ints -> (16,) uint64: x.c0 x.c1 y.c0 y.c1 in Montgomery form (arkworks' G2Affine)."""
    if isinstance(pt, np.ndarray):                       # already in the ABI's form (e.g. from keyfile.cvk_read)
        return np.ascontiguousarray(pt, dtype=np.uint64).reshape(16)
    out = np.zeros(16, dtype=np.uint64)
    if pt is not None:
        (x0, x1), (y0, y1) = pt
        for k, v in enumerate((x0, x1, y0, y1)):
            out[4 * k: 4 * k + 4] = _fq_mont_limbs(v)
    return out


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# alt_bn128 G2 generator (EIP-197), ((x.c0, x.c1), (y.c0, y.c1))
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


def g2_mul(k, pt=G2_GEN):
    """k * pt on G2 through the C ABI (zkb_g2_mul); returns the (16,) Montgomery array the verifier takes."""
    k = int(k)
    sc = np.array([(k >> (64 * j)) & (2**64 - 1) for j in range(4)], dtype=np.uint64)
    out = np.zeros(16, dtype=np.uint64)
    rc = _lib.lib().zkb_g2_mul(_vp(g2_array(pt)), _vp(sc), _vp(out))
    if rc != 0:
        raise _lib.ZkbError(rc, "zkb_g2_mul: point not on the twist")
    return out


def make_cvk(tau):
    """(h, beta_h) = (H, tau * H) for a synthetic SRS with known trapdoor (kzg10::setup draws h at random; any h gives the
    same accept / reject decisions).  Product code: bench.py verifies every timed proof with it."""
    return g2_array(G2_GEN), g2_mul(tau)


def verify(vk, proof_bytes, pub_inputs, cvk_g2, transcript="merlin"):
    """vk: prover.VerifierKey; proof_bytes: the 802 serialised bytes; pub_inputs: canonical ints, one per vk.pi_roots entry;
    cvk_g2 = (h, beta_h) as Fq2 coordinate tuples of canonical ints or as the (16,) arrays keyfile.cvk_read returns.  Returns 0 (accepted), 1 or 2 (failing step); raises on malformed input."""
    assert len(proof_bytes) == 802 and len(pub_inputs) == len(vk.pi_roots), "invalid length of public inputs"
    xy = np.stack([g1_array(vk.commits[name]) for name in VerifierKey.ORDER])
    inf = (ctypes.c_int * 10)(*[int(vk.commits[name] is None) for name in VerifierKey.ORDER])
    roots = ints_to_mont_array(vk.pi_roots) if vk.pi_roots else np.zeros((1, 4), dtype=np.uint64)
    pub = ints_to_mont_array(list(pub_inputs)) if pub_inputs else np.zeros((1, 4), dtype=np.uint64)
    raw = np.frombuffer(bytes(proof_bytes), dtype=np.uint8).copy()
    h, bh = g2_array(cvk_g2[0]), g2_array(cvk_g2[1])
    rc = _lib.lib().zkb_plonk_verify(vk.n, _vp(roots), len(vk.pi_roots), _vp(xy), inf, _vp(pub), _vp(raw), _vp(h), _vp(bh),
                                     TRANSCRIPTS[transcript][0])
    if rc < 0:
        raise _lib.ZkbError(rc, "zkb_plonk_verify: malformed verifier key, proof or G2 elements")
    return rc


def pairing(p1, q2):
    """e(P, Q): 12 canonical Fq12 coefficients in the w-basis (ints)."""
    out = np.zeros((12, 4), dtype=np.uint64)
    rc = _lib.lib().zkb_pairing(_vp(g1_array(p1)), _vp(g2_array(q2)), _vp(out))
    if rc != 0:
        raise _lib.ZkbError(rc, "zkb_pairing: point not on its curve")
    return [sum(int(out[k, j]) << (64 * j) for j in range(4)) for k in range(12)]


def pairing_product_is_one(pairs):
    g1 = np.stack([g1_array(p) for p, _ in pairs])
    g2 = np.stack([g2_array(q) for _, q in pairs])
    one = ctypes.c_int(0)
    rc = _lib.lib().zkb_pairing_product_is_one(_vp(g1), _vp(g2), len(pairs), ctypes.byref(one))
    if rc != 0:
        raise _lib.ZkbError(rc, "zkb_pairing_product_is_one: point not on its curve")
    return bool(one.value)
