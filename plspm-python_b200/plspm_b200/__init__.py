"""plspm_b200: engine binding of the B200-native PLS-PM weight-estimation library."""
