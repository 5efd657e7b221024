"""hig_b200 — B200-native (sm_100a) denoising hot path of line/Human-Interaction-Generation.

The directory is named after the reference repo (``human-interaction-generation_b200``), which is not a valid
Python identifier; import it through the root-level shim module ``hig_b200`` (``import hig_b200``).

Layout mirrors the reference's ``codes/`` files on the hot path:
  interaction_transformer.py  <- codes/models/interaction_transformer.py (MotionInteractionTransformer)
  gaussian_diffusion.py       <- codes/models/gaussian_diffusion.py     (GaussianDiffusion)
  mul_ddpm_trainer.py         <- codes/trainers/mul_ddpm_trainer.py     (DDPMMulTrainer)
  ops.py / _lib.py            tensor wrappers + ctypes binding of the C ABI (include/hig_b200.h)
  csrc/                       hand-written CUDA (tcgen05/TMEM/TMA GEMM, fused attention, LN+FiLM, DDPM step)
"""
__version__ = "0.1.0"
