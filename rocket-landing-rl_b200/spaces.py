"""``Box`` space with the attributes the reference's callers read
(``backend/envs/lander.py:65-69,103-107``; ``backend/rl/agent.py`` reads ``.shape``).
``gymnasium.spaces.Box`` is used when gymnasium is importable, this stand-in otherwise."""
from __future__ import annotations

import numpy as np

try:                                       # pragma: no cover - gymnasium is absent in the image
    from gymnasium.spaces import Box       # type: ignore
except Exception:                          # noqa: BLE001
    class Box:                             # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            low = np.asarray(low, dtype=self.dtype)
            high = np.asarray(high, dtype=self.dtype)
            if shape is not None:
                low = np.broadcast_to(low, shape).copy()
                high = np.broadcast_to(high, shape).copy()
            self.low, self.high = low, high
            self.shape = tuple(low.shape)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return (x.shape == self.shape and bool(np.all(x >= self.low))
                    and bool(np.all(x <= self.high)))

        def __contains__(self, x) -> bool:
            return self.contains(x)

        def __eq__(self, other) -> bool:
            return (isinstance(other, Box) and self.shape == other.shape
                    and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

        def __repr__(self) -> str:
            return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"
