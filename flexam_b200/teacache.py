"""TeaCache step-skipping state (semantics of FlexAM/models/cache_utils.py:21-76 and its use in the reference
forward, wan_transformer3d_FlexAM.py:977-1051, :1119-1122), kept as plain host-side control flow around the native
block loop: a polynomial-rescaled relative-L1 distance of the modulated timestep embedding accumulates until it
crosses ``rel_l1_thresh``; below it the 30 blocks are skipped and the previous residual (x_after - x_before) is
re-applied with one fused add."""
from __future__ import annotations

from typing import List, Optional

import torch


class TeaCache:
    def __init__(self, coefficients: List[float], num_steps: int, rel_l1_thresh: float = 0.0,
                 num_skip_start_steps: int = 0, offload: bool = True):
        if num_steps < 1:
            raise ValueError(f"`num_steps` must be greater than 0 but is {num_steps}.")
        if rel_l1_thresh < 0:
            raise ValueError(f"`rel_l1_thresh` must be greater than or equal to 0 but is {rel_l1_thresh}.")
        if num_skip_start_steps < 0 or num_skip_start_steps > num_steps:
            raise ValueError("`num_skip_start_steps` must be in [0, num_steps]")
        self.coefficients = list(coefficients)
        self.num_steps = num_steps
        self.rel_l1_thresh = rel_l1_thresh
        self.num_skip_start_steps = num_skip_start_steps
        self.offload = offload          # accepted for API parity; residuals stay in HBM (180 GB) on this path
        self.reset()

    def rescale_func(self, x: float) -> float:
        y = 0.0
        for c in self.coefficients:     # np.poly1d order: highest power first
            y = y * x + c
        return y

    @staticmethod
    def compute_rel_l1_distance(prev: torch.Tensor, cur: torch.Tensor) -> float:
        return ((cur - prev).abs().mean() / prev.abs().mean()).item()

    def reset(self):
        self.cnt = 0
        self.should_calc = True
        self.accumulated_rel_l1_distance = 0.0
        self.previous_modulated_input: Optional[torch.Tensor] = None
        self.previous_residual = None
        self.previous_residual_cond = None
        self.previous_residual_uncond = None

    def decide(self, modulated_inp: torch.Tensor, cond_flag: bool) -> bool:
        return decide(self, modulated_inp, cond_flag)

    def step_done(self, cond_flag: bool):
        step_done(self, cond_flag)


# The control flow of the reference forward around its TeaCache object, written against the object's ATTRIBUTES so
# that it drives this module's TeaCache and the reference's own class (FlexAM/models/cache_utils.py:21-76, which has the
# same state and no decide()/step_done()) alike: an installed reference module whose pipeline called the reference's
# enable_teacache() keeps working.
def decide(tc, modulated_inp: torch.Tensor, cond_flag: bool) -> bool:
    """Decision for one forward call (:978-1000)."""
    if not cond_flag:
        return tc.should_calc
    if tc.cnt < tc.num_skip_start_steps:
        should = True
        tc.accumulated_rel_l1_distance = 0.0
    else:
        d = tc.compute_rel_l1_distance(tc.previous_modulated_input, modulated_inp)
        tc.accumulated_rel_l1_distance += float(tc.rescale_func(d))
        if tc.accumulated_rel_l1_distance < tc.rel_l1_thresh:
            should = False
        else:
            should = True
            tc.accumulated_rel_l1_distance = 0.0
    tc.previous_modulated_input = modulated_inp.clone()
    tc.should_calc = should
    return should


def step_done(tc, cond_flag: bool) -> None:
    if cond_flag:                       # :1119-1122
        tc.cnt += 1
        if tc.cnt == tc.num_steps:
            tc.reset()
