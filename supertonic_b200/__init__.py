"""supertonic_b200 — B200-native (sm_100a) synthesis forward pass for Supertonic TTS.

Only what the hot path needs: the C-ABI CUDA library under ``csrc/``
(`libsupertonic_cuda.so`), its ctypes binding (`capi`), the host-side mirror of the
reference `TextToSpeech` API (`tts`), the utterance scheduler (`scheduler`) and the
surrogate asset generator (`surrogate`, because the released weights are not mounted).
There is no CPU fallback for the neural path: importing `capi` without the built
library raises.
"""
__version__ = "0.1.0"
