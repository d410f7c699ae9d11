"""llmvox_b200: B200-native (sm_100a) implementation of LLMVoX's speech-synthesis hot path.

Device work lives in libllmvox_b200.so (hand-written CUDA behind the C ABI of include/llmvox_b200.h); this
package is the host-side mirror of the reference's Python interface for the path.  Importing the package does
not need a GPU; constructing an Engine / ModelHandler does, and there is no CPU fallback."""

__all__ = ["weights", "tokenizer", "scheduler"]
