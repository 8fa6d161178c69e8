"""Test-side stand-in for era_boojum_b200.parallel.TorchBackend: the same interface on torch CPU int64 tensors, computed by the
oracle, so that the sharding / gathering logic of parallel.commit_sharded can run under gloo without a GPU.  Lives in tests/
because only tests may touch oracle/."""
import numpy as np


class OracleBackend:
    def __init__(self):
        import torch
        from oracle import oracle as O
        self.torch, self.O = torch, O

    def _np(self, t):
        return t.numpy().view(np.uint64)

    def _t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a).view(np.int64))

    def intt(self, cols):
        return self._t(self.O.intt_n2n(self._np(cols)))

    def coset_ntt(self, monomials, shift):
        return self._t(self.O.ntt_n2b(self._np(monomials), shift))

    def subtree(self, cols_2d, cap):
        a = self._np(cols_2d)
        lh, levels, capd = self.O.merkle_tree([a[c] for c in range(a.shape[0])], cap)
        return (lh, levels), self._t(capd)

    def empty(self, shape, like):
        return self.torch.empty(shape, dtype=self.torch.int64)
