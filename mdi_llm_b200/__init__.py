"""B200-native model-distributed inference (MDI) engine with the capabilities of MDI-LLM."""
__version__ = "0.1.0"
