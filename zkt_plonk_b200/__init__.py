"""zkt_plonk_b200 -- B200-native (sm_100a) prover hot path of ZKTLabs/zkt-plonk.

Host-side mirror of the reference's generic seams over libzkb200.so (include/zkb200.h):
  GpuEvaluationDomain  <-> D: EvaluationDomain + EvaluationDomainExt   (plonk-core/src/util.rs:27-140)
  GpuKZG10             <-> PC: HomomorphicCommitment (KZG10<Bn254>)     (plonk-core/src/commitment.rs:10-46)
The CUDA library is mandatory: importing a wrapper without libzkb200.so raises ImportError.
"""
from ._lib import LIB_PATH, ZkbError, declared_symbols  # noqa: F401
from .context import Context, sum_partials  # noqa: F401
from .domain import GpuEvaluationDomain  # noqa: F401
from .kzg import GpuKZG10, PCError  # noqa: F401
from . import keyfile, prover_ops  # noqa: F401,E402
# `prover` (round schedule) and `verifier` (zkb_plonk_verify) are imported on demand: they pull in the transcript code
