"""mfas_b200 -- B200-native candidate-training hot path of MFAS."""
__version__ = "0.1.0"
