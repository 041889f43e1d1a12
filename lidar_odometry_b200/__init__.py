"""lidar_odometry_b200 — B200-native scan-to-map registration engine (hot path of SiarheiHerasiuta/lidar_odometry).

``capi``  ctypes binding of libb2lo.so (C ABI of include/b2lo.h; hand-written sm_100a kernels in csrc/)
``api``   host-side mirror of the reference's FastVoxelFilter / VoxelMap / IterativeClosestPointOptimizer classes
``synth`` seeded synthetic KITTI- and MID360-shaped sequences
``shim/`` C++ drop-in classes with the reference's own names over the same C ABI

No CPU fallback: importing ``api`` works anywhere, but every call needs libb2lo.so and a CUDA device.
"""
__version__ = "0.1.0"
