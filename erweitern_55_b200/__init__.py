"""B200-native drop-in for erweitern_55's batched policy/value network evaluation."""
