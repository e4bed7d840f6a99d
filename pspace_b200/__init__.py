"""pspace_b200 — B200-native stochastic Galerkin projection behind the pspace ParameterContainer API.

  cabi      ctypes binding of include/pspace_cuda.h (lib/libpspace_cuda.so: the CUDA kernels)
  device    DeviceContainer: device-resident container driven with torch tensors (tests, bench, multi-GPU)
  PSPACE    PyParameterFactory / PyParameterContainer over lib/libpspace.so (the reference wrapper's API)
  sharded   quadrature-point sharding across ranks (torch.distributed)
  workloads the five BASELINE.json configurations
"""
__all__ = ["cabi", "device", "PSPACE", "cwrap", "sharded", "workloads"]
