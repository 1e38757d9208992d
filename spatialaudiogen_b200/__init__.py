"""spatialaudiogen_b200: B200-native inference hot path of spatialaudiogen."""
