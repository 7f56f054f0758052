"""B200-native (sm_100a) implementation of the STEM P-frame hot path behind the CompressAI model API."""
__version__ = "0.2.0"
