"""B200-native PointPillars pre/post-processing hot path."""
