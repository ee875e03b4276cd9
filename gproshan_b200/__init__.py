"""B200-native Parallel Toplesets Propagation (PTP) geodesics — drop-in for larc/gproshan's PTP_GPU path."""
