"""Drop-in for `data_utils.KernelSmoothing` (neuripss2020/kccotgan, data_utils.py:478-586).

Only the kernel-smoothing class is on the loss hot path; the loaders, plotting helpers and LR
schedules of the reference's data_utils are out of scope.
"""
import numpy as np
import torch

from .functional import SmoothFn, _check


class KernelSmoothing:
    """Gaussian smoothing of [B,H,T,W,C] videos, divided by the global max of the result."""

    def __init__(self, temporal_kernel_size=6, spatial_kernel_size=8):
        # data_utils.py:479-481
        self.temporal_radius = temporal_kernel_size // 2
        self.spatial_radius = spatial_kernel_size // 2
        self._filters = {}

    # -- weights --------------------------------------------------------------------------------
    @staticmethod
    def _weights(radius, sigma):
        """fp32 arithmetic exactly as data_utils.py:487-490."""
        sigma2 = np.float32(sigma) * np.float32(sigma)
        x = np.arange(-radius, radius + 1, dtype=np.float32)
        k = np.exp((np.float32(-0.5) / sigma2) * x ** 2, dtype=np.float32)
        return (k / k.sum(dtype=np.float32)).astype(np.float32)

    def gaussian_kernel1d(self, radius, sigma, device="cuda"):
        """data_utils.py:483-491 — [2r+1] normalised Gaussian weights."""
        return torch.from_numpy(self._weights(radius, sigma)).to(device)

    def gaussian_kernel3d(self, radius, sigma, device="cuda"):
        """data_utils.py:493-501 — [2r+1,2r+1,2r+1,1,1]; equals w (x) w (x) w up to fp32 rounding."""
        x = np.arange(-radius, radius + 1, dtype=np.float32)
        xx, yy, zz = np.meshgrid(x, x, x)
        sigma2 = np.float32(sigma) * np.float32(sigma)
        k = np.exp((np.float32(-0.5) / sigma2) * (xx ** 2 + yy ** 2 + zz ** 2), dtype=np.float32)
        k = (k / k.sum(dtype=np.float32)).astype(np.float32)
        return torch.from_numpy(k[:, :, :, None, None]).to(device)

    def _filter_matrix(self, n, radius, sigma, device):
        """Dense [n,n] matrix of REFLECT pad (data_utils.py:513) + VALID cross-correlation (:515)."""
        key = (n, radius, float(sigma), str(device))
        if key not in self._filters:
            if n <= radius:
                raise ValueError(f"REFLECT padding by {radius} needs an axis longer than {radius}, got {n}")
            if radius != 3:
                raise ValueError("libkccot's smoothing kernels are built for kernel size 6 (radius 3), the value "
                                 "kernel_train.py:216 uses")
            w = self._weights(radius, sigma)
            A = np.zeros((n, n), dtype=np.float32)
            for p in range(n):
                for k in range(2 * radius + 1):
                    q = p + k - radius
                    q = -q if q < 0 else q
                    q = 2 * (n - 1) - q if q >= n else q
                    A[p, q] += w[k]
            if len(self._filters) > 64:
                self._filters.clear()
            self._filters[key] = torch.from_numpy(A).to(device)
        return self._filters[key]

    # -- convolutions ---------------------------------------------------------------------------
    def temporal_convolution(self, inputs, sigma):
        """data_utils.py:503-521 — 7-tap REFLECT filter along T, / global max."""
        inputs = _check(inputs, "inputs", 5)
        ft = self._filter_matrix(inputs.shape[2], self.temporal_radius, sigma, inputs.device)
        return SmoothFn.apply(inputs, 1, None, ft, None)

    def spatial_convolution(self, inputs, sigma):
        """data_utils.py:523-550 — BROKEN in the reference: the VALID conv2d shrinks H, W by 2r and the
        following reshape to the original [.., h, w] fails (:537-538, :547-548).  Reproduced as an error."""
        bs, h, t, w, nc = inputs.shape
        r = self.spatial_radius
        raise ValueError(f"spatial_convolution: cannot reshape the VALID-convolved tensor of "
                         f"{bs * nc * t * (h - 2 * r) * (w - 2 * r)} elements into {[bs, nc, t, h, w]} "
                         "(the reference's '2d' kernel raises here, data_utils.py:537-538)")

    def gaussian_convolution3D(self, inputs, sigma):
        """data_utils.py:552-582 — 7^3 REFLECT filter over (H,T,W) per channel (three separable
        7-tap passes), / global max.  Uses the SPATIAL radius for all three axes (:553,562-564)."""
        inputs = _check(inputs, "inputs", 5)
        _, H, T, W, _ = inputs.shape
        r, dev = self.spatial_radius, inputs.device
        return SmoothFn.apply(inputs, 3, self._filter_matrix(H, r, sigma, dev), self._filter_matrix(T, r, sigma, dev),
                              self._filter_matrix(W, r, sigma, dev))

    def annealing_sigma(self, init_sigma, step, decay_steps=500, decay_rate=0.975):
        """data_utils.py:584-586."""
        return init_sigma * decay_rate ** (step / decay_steps)
