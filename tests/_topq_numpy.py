"""TEST DOUBLE: numpy implementation of the local steps of the radix top-q (same digit layout as
csrc/sampler.cu) so that the distributed protocol in sgs_gnn_b200/dist.py can run under gloo on CPU."""
import numpy as np
import torch

SHIFT = (20, 9, 0)
BINS = (2048, 2048, 512)


class NumpyTopQOps:
    def keys(self, p, prob, noise, mode, coef, S):
        p32 = p.numpy().astype(np.float32)
        if mode == 2:
            s = p32
        else:
            s_eff = np.float32(S.numpy()[0]) + np.float32(1e-12)
            s = p32 / s_eff
            if mode == 0:
                s = np.float32(1.0 - coef) * s + np.float32(coef) * prob.numpy().astype(np.float32)
        key = (s / noise.numpy().astype(np.float32)).astype(np.float32)
        bits = key.view(np.uint32) & np.uint32(0x7FFFFFFF)
        hist = torch.from_numpy(np.bincount(bits >> np.uint32(20), minlength=2048).astype(np.int64))
        return torch.from_numpy(bits.astype(np.int64)), hist, torch.zeros(8, dtype=torch.int64)

    def hist(self, keys, hist, state, level):
        k = keys.numpy().astype(np.uint32)
        hi_shift = 20 if level == 1 else 9
        match = (k >> np.uint32(hi_shift)) == np.uint32(int(state[0]) >> hi_shift)
        digit = (k[match] >> np.uint32(SHIFT[level])) & np.uint32(BINS[level] - 1)
        hist.copy_(torch.from_numpy(np.bincount(digit, minlength=2048).astype(np.int64)))

    def find(self, hist, state, k_total, level):
        k = k_total if level == 0 else int(state[1])
        h = hist.numpy()
        cum, b = 0, BINS[level] - 1
        while b > 0:
            if cum + h[b] >= k:
                break
            cum += h[b]
            b -= 1
        prefix = (0 if level == 0 else int(state[0])) | (b << SHIFT[level])
        state[0], state[1] = prefix, k - cum
        if level == 2:
            state[2], state[3], state[4], state[6] = prefix, k_total - (k - cum), k - cum, int(h[b])
        hist.zero_()

    def compact(self, keys, state, tie_skip, q_cap, n_expected=None):
        k = keys.numpy().astype(np.uint32)
        tau = np.uint32(int(state[2]))
        avail = max(int(state[4]) - tie_skip, 0)
        gt = np.nonzero(k > tau)[0]
        eq = np.nonzero(k == tau)[0][:avail]
        sel = np.sort(np.concatenate([gt, eq])).astype(np.int32)
        assert n_expected is None or n_expected == sel.size, (n_expected, sel.size)
        return torch.from_numpy(sel)
