"""cosdata_b200 -- B200-native (sm_100a) implementation of cosdata's ANN distance hot path.

The package is a thin host mirror of the reference's operator surface over the
C ABI in include/cosdata_b200.h; all compute lives in csrc/*.cu.
"""
from .api import (  # noqa: F401
    INVALID_ID, CosdataError, DenseIndex, HnswFiles, DistanceError, DistanceMetric, DistanceMetricKind,
    ScalarQuantization, SearchMode, Status, Storage, StorageType, code_bytes, debug_set_hnsw_flags, device_count,
    itoe_get, itoe_load, itoe_scan, kernel_launch_count, prop_file_load, prop_file_load_metadata, prop_file_scan, sample_values_range, sample_values_range_device, synth_matrix, tensor_peak,
)
