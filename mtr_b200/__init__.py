"""mtr_b200 -- B200-native (sm_100a) implementation of mTR's per-read tandem-repeat hot path."""
