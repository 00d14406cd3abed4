"""fv2d_b200 — B200-native (sm_100a) implementation of the per-timestep finite-volume update
of mdelorme/fv2d, behind a C ABI (include/fv2d_b200.h).

  fv2d_b200.capi      ctypes view of the C ABI (Context, params_from_ini, init_problem)
  fv2d_b200.multigpu  one-process-per-GPU y-slab plumbing over torch.distributed
  fv2d_b200/csrc      CUDA kernels + C ABI implementation
  fv2d_b200/host      C++17 host mirror of the reference's operator surface + driver
"""
