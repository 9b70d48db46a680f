"""B200-native hot path of chenglou/ai-font-renderer (training step + batched render)."""
__all__ = ["build"]
