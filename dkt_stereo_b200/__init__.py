"""B200-native inference engine for the DKT-Stereo hot path."""
__version__ = "0.1.0"
