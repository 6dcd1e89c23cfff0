"""fugue_b200 - a B200-native (sm_100a) implementation of Fugue's
``fa.transform() -> MapEngine.map_dataframe`` hot path (hash PartitionSpec),
plus ``ExecutionEngine.join`` / ``aggregate`` kernels.

The CUDA library (``libfugue_b200.so``, C ABI in ``include/fugue_b200.h``) is the
product; this package is the thin host-side mirror of the reference's plugin
interface.  There is no CPU fallback on any engine path.
"""
__version__ = "0.1.0"
