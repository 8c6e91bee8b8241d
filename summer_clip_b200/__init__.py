"""summer-clip_b200: B200-native CLIP-search hot path.

Importing the package does not touch the GPU; the C-ABI library is loaded on first use and its
absence is an error (there is no CPU / PyTorch fallback on the product path)."""
__version__ = "0.1.0"
