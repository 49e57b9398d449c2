"""Module plumbing shared by the index and its codecs.

Checkpoint compatibility with the reference: its objects are ``nn.Module``s whose state is
a set of registered buffers, and its ``load_state_dict`` drops and re-registers every
top-level buffer so shapes may change between save and load
(torchpq/CustomModule.py:14-23; README.md:90-97).  ``StateModule`` keeps that contract and
the same buffer names, so a ``state_dict`` saved by ``torchpq.index.IVFPQIndex`` loads here.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class StateModule(nn.Module):
    verbose = 0

    def print_message(self, text, min_verbosity=0):
        if self.verbose >= min_verbosity:
            print(f"{type(self).__name__}: {text}")

    def load_state_dict(self, state_dict, strict=True):  # noqa: D401 - reference semantics, see module docstring
        for key, value in state_dict.items():
            if "." in key:
                continue
            assert hasattr(self, key), f"attribute {key} does not exist"
            delattr(self, key)
            self.register_buffer(key, value)
        for name, child in self.named_children():
            prefix = name + "."
            child.load_state_dict({k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)})
        self._state_changed()

    def _state_changed(self):
        pass


class _Centroids(StateModule):
    """Stands where the reference has KMeans / MultiKMeans: owner of the ``centroids`` buffer
    (clustering/KMeans.py:74, clustering/MultiKMeans.py:73)."""

    def __init__(self):
        super().__init__()
        self.register_buffer("centroids", None)


class Codec(StateModule):
    """BaseCodec (codec/BaseCodec.py:5-17): ``_is_trained`` flag + a ``kmeans.centroids`` codebook."""

    def __init__(self):
        super().__init__()
        self.register_buffer("_is_trained", torch.tensor(False))
        self.kmeans = _Centroids()

    @property
    def is_trained(self) -> bool:
        # cached as a Python bool: `.item()` on a CUDA `_is_trained` would be a host sync on every search
        t = self._is_trained
        key = (t.data_ptr(), t._version)
        if getattr(self, "_trained_key", None) != key:
            self._trained_cache, self._trained_key = bool(t.item()), key
        return self._trained_cache

    @property
    def codebook(self):
        return self.kmeans.centroids if self.is_trained else None

    def set_codebook(self, centroids: torch.Tensor):
        self.kmeans.register_buffer("centroids", centroids)
        self._is_trained.data = torch.tensor(True)
