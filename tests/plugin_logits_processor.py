"""Test plug-in for ``--logits-processors``: forces the token id given in ``extra_args["force_token"]``."""
import torch

from vllm_b200.sample.sampler import LogitsProcessor


class ForceToken(LogitsProcessor):
    def apply(self, logits: torch.Tensor, md) -> torch.Tensor:
        for row, ea in enumerate(md.extra_args):
            if ea and "force_token" in ea:
                keep = logits[row, ea["force_token"]].clone()
                logits[row] = float("-inf")
                logits[row, ea["force_token"]] = keep
        return logits
