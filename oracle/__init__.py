"""CPU oracle: a restatement of the reference's sampling / quantize / decode path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker (or as the timed CPU
baseline), never as the thing shipped.  The product package fails loudly when its CUDA
library is missing; it never routes through this package.

PARITY UNPINNED: the reference (aayush9400/3D-Condtional-Stable-Diffusion) is Python on
TensorFlow 2.12 / Keras 2.  TensorFlow is not installed in this environment and cannot be
(no network), and the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md F2/F3, section 8c).  The restatement below therefore follows the reference *source*
(every function cites the file:line it restates) and the documented Keras op semantics;
the only reference-derived pins are the Keras parameter count logged in
``experiments/vqvae3d-scaled-monai-B8-AUG-all-T-KR.output:23-25`` (reproduced from the encoder + decoder weight tables)
and the closed-form schedule identities of ``Betas`` (both in tests/test_oracle_cpu.py), next to the Random123
known-answer vectors for Philox4x32-10.

Arithmetic: PyTorch-CPU fp32 (``dtype=torch.float64`` for an independent high-precision
check).  ``Emu`` (oracle.ops) optionally rounds to bf16 at exactly the points where the
CUDA path stores bf16, so that the kernel-vs-oracle comparison isolates kernel error
from the (stated) bf16 storage error.
"""
from . import ops, schedule, unet, first_stage, sampler, philox, init  # noqa: F401
