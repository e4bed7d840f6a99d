"""Synthetic workloads = the five configurations of BASELINE.json (sizes table and §8d).

A workload is only a description (parameter list, basis type, block edge m, seed); it carries no
arithmetic.  `params` entries are (kind, p1, p2, dmax) in ParameterFactory order
(createNormalParameter(mu, sigma, dmax) / createUniformParameter(a, b, dmax) /
createExponentialParameter(mu, beta, dmax), reference src/include/ParameterFactory.h:26-28).
Quadrature is always the reference default `initialize()`: nq_d = dmax_d + 1
(src/cpp/ParameterContainer.cpp:224-237).
"""
from dataclasses import dataclass, field
from math import comb, prod
from typing import List, Tuple

TENSOR, COMPLETE = 0, 1


def mixed_params(nvars: int, dmax: int) -> List[Tuple[str, float, float, int]]:
    """Normal(1+0.1d, 0.1), Uniform(-1, 1+0.5d), Exponential(1, 0.1+0.05d) cycled by index d."""
    out = []
    for d in range(nvars):
        if d % 3 == 0:
            out.append(("normal", 1.0 + 0.1 * d, 0.1, dmax))
        elif d % 3 == 1:
            out.append(("uniform", -1.0, 1.0 + 0.5 * d, dmax))
        else:
            out.append(("exponential", 1.0, 0.1 + 0.05 * d, dmax))
    return out


@dataclass
class Workload:
    name: str
    cfg: int
    params: list
    basis_type: int
    m: int                      # block edge: J_q is m x m (m2 = m*m columns), f_q has m entries
    complex_: bool = False
    elements: int = 1           # cfg 3: elements sharing one container
    note: str = ""
    pair_window: Tuple[int, int] = field(default=None)  # (rows i, cols j) when the full N^2 is not materialisable

    @property
    def seed(self) -> int:
        return 20260 + self.cfg

    @property
    def nvars(self) -> int:
        return len(self.params)

    @property
    def Q(self) -> int:
        return prod(p[3] + 1 for p in self.params)

    @property
    def N(self) -> int:
        dm = [p[3] for p in self.params]
        if self.basis_type == TENSOR or self.nvars == 1:
            return prod(d + 1 for d in dm)
        if len(set(dm)) == 1:
            return comb(self.nvars + dm[0], dm[0])
        raise ValueError("closed form only for equal dmax; ask the container")

    @property
    def m2(self) -> int:
        return self.m * self.m


def cfg1():
    return Workload("cfg1_3param_tensor_p3_scalar", 1,
                    [("normal", 1.0, 0.1, 3), ("uniform", 1.0, 2.0, 3), ("exponential", 1.0, 0.1, 3)],
                    TENSOR, 1, note="test_pspace.py scale; KATs of SURVEY.md §8c")


def cfg2():
    return Workload("cfg2_5param_complete_p4_m8", 2, mixed_params(5, 4), COMPLETE, 8)


def cfg3():
    return Workload("cfg3_4param_tensor_p5_m24", 3, mixed_params(4, 5), TENSOR, 24, elements=10000,
                    note="beam-example scale; 7.7 GB of output per element -> benchmarked E' = 8 elements per step through the "
                         "batched-elements call (full 1296 x 1296 window, 62 GB of output), per-element rate extrapolated to 10k")


def cfg4():
    ps = [("normal", float(d), 1.0, 3) for d in range(8)]
    return Workload("cfg4_8gauss_tensor_p3_m24_complex", 4, ps, TENSOR, 24, complex_=True,
                    note="full N^2 output is 39.6 PB -> matrix projection on a pair window",
                    pair_window=(64, 64))


def cfg5():
    return Workload("cfg5_10param_complete_p3_m24", 5, mixed_params(10, 3), COMPLETE, 24)


ALL = {1: cfg1, 2: cfg2, 3: cfg3, 4: cfg4, 5: cfg5}
