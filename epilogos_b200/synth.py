"""Synthetic state matrices of the benchmark shapes, generated on the device.

"realistic": per-bin dominant state d_b ~ prior, each label = d_b with probability 0.6, else ~ prior; the
prior is skewed to the last (quiescent) state (~62 %).  "uniform": iid uniform labels (adversarial for S2:
every state present in every bin).  torch is used here only to make test / benchmark INPUT data.
"""
import torch

from .engine import pitch_for


def state_prior(num_states):
    w = torch.tensor([0.5 ** (i * 0.6) for i in range(num_states - 1)], dtype=torch.float64)
    w = 0.38 * w / w.sum()
    return torch.cat([w, torch.tensor([0.62], dtype=torch.float64)])


def synth_states_device(bins, cols, num_states, seed, kind="realistic", device="cuda", chunk=1 << 18):
    """Returns a CUDA int8 tensor [bins, pitch_for(cols)]; columns >= cols are zero padding."""
    pitch = pitch_for(cols)
    out = torch.zeros((bins, pitch), dtype=torch.int8, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    prior = state_prior(num_states).to(device=device, dtype=torch.float32)
    cdf = torch.cumsum(prior, 0)
    cdf[-1] = 2.0
    for lo in range(0, bins, chunk):
        n = min(chunk, bins - lo)
        if kind == "uniform":
            lab = torch.randint(0, num_states, (n, cols), generator=gen, device=device, dtype=torch.int8)
        else:
            dom = torch.bucketize(torch.rand((n, 1), generator=gen, device=device), cdf).to(torch.int8)
            other = torch.bucketize(torch.rand((n, cols), generator=gen, device=device), cdf).to(torch.int8)
            keep = torch.rand((n, cols), generator=gen, device=device) < 0.6
            lab = torch.where(keep, dom.expand(n, cols), other)
            del dom, other, keep
        out[lo:lo + n, :cols] = lab
        del lab
    return out
