"""Seeded synthetic k-SAT instance generators and DIMACS text I/O helpers.

The reference ships no instances; these follow the recipe of
SURVEY.md section 8d: *configuration model* -- ``d`` slots per variable, shuffle, cut into
k-tuples, drop tuples with a repeated variable, fair random signs -- and uniform
random k-SAT.  Literal encoding is the reference's (example/main.cpp:168):
``lit = 2*var + neg`` with 0-based ``var``; DIMACS ``x>0 -> 2x-2``, ``-x -> 2x-1``.

numpy versions (host, any size that fits RAM) and torch versions (generated
directly in HBM for the large bench configurations).
"""
from __future__ import annotations

import numpy as np

#: BASELINE.json configs -> (k, n_vars, d) for the bounded-degree shapes
CONFIGS = {
    "cfg1": dict(kind="bounded", k=5, n=2_000, d=3),          # m ~ 1,200 (DIMACS, reference serial)
    "cfg2": dict(kind="bounded", k=7, n=1_000_000, d=28),     # m ~ 4 M
    "cfg3": dict(kind="uniform", k=3, n=1_000_000, m=3_000_000),
    "cfg4": dict(kind="bounded", k=8, n=10_000_000, d=32),    # m ~ 40 M
    "cfg5": dict(kind="bounded", k=5, n=10_000, d=3),         # m ~ 6,000 per instance, x 8,192
}
INSTANCE_SEED_BASE = 0xA111


def bounded_degree_ksat(n: int, k: int, d: int, seed: int) -> np.ndarray:
    """Every variable occurs at most ``d`` times.  Returns an (m, k) uint32 literal matrix."""
    rng = np.random.default_rng(seed)
    slots = np.repeat(np.arange(n, dtype=np.uint32), d)
    rng.shuffle(slots)
    m = len(slots) // k
    tuples = slots[: m * k].reshape(m, k)
    srt = np.sort(tuples, axis=1)
    ok = (srt[:, 1:] != srt[:, :-1]).all(axis=1)
    tuples = tuples[ok]
    signs = rng.integers(0, 2, size=tuples.shape, dtype=np.uint32)
    return np.ascontiguousarray(tuples * np.uint32(2) + signs, dtype=np.uint32)


def uniform_ksat(n: int, k: int, m: int, seed: int) -> np.ndarray:
    """Uniform random k-SAT: k distinct variables per clause, no occurrence bound."""
    rng = np.random.default_rng(seed)
    vars_ = rng.integers(0, n, size=(m, k), dtype=np.int64)
    while True:  # redraw the (rare) clauses with a repeated variable
        srt = np.sort(vars_, axis=1)
        bad = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
        nb = int(bad.sum())
        if nb == 0:
            break
        vars_[bad] = rng.integers(0, n, size=(nb, k), dtype=np.int64)
    signs = rng.integers(0, 2, size=(m, k), dtype=np.int64)
    return np.ascontiguousarray(vars_ * 2 + signs, dtype=np.uint32)


def make_config(name: str, scale: float = 1.0, seed: int | None = None) -> tuple[int, np.ndarray]:
    """(n_vars, lits[m,k]) for a BASELINE.json config, optionally scaled down in n."""
    cfg = CONFIGS[name]
    idx = int(name[3:])
    seed = INSTANCE_SEED_BASE + idx if seed is None else seed
    n = max(int(cfg["n"] * scale), cfg["k"] * 4)
    if cfg["kind"] == "bounded":
        return n, bounded_degree_ksat(n, cfg["k"], cfg["d"], seed)
    return n, uniform_ksat(n, cfg["k"], max(int(cfg["m"] * scale), 1), seed)


def bounded_degree_ksat_torch(n: int, k: int, d: int, seed: int, device="cuda"):
    """Same recipe generated in device memory; returns an (m, k) int32-viewed-uint32 torch tensor.

    (A different random stream from the numpy version: instances are reproducible per
    (seed, generator), not across generators.)
    """
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    total = n * d
    m = total // k
    perm = torch.randperm(total, generator=g, device=device)[: m * k]
    tuples = (perm // d).to(torch.int32).reshape(m, k)
    del perm
    srt, _ = torch.sort(tuples, dim=1)
    ok = (srt[:, 1:] != srt[:, :-1]).all(dim=1)
    del srt
    tuples = tuples[ok]
    signs = torch.randint(0, 2, tuples.shape, generator=g, device=device, dtype=torch.int32)
    return (tuples * 2 + signs).contiguous()


def bounded_degree_batch_torch(n_inst: int, n: int, k: int, d: int, seed: int, device="cuda"):
    """``n_inst`` independent bounded-degree instances generated on the device in one go (BASELINE config 5).

    Returns (clause_off int64 [n_inst+1] on the host, lits int32 [total, k] on the device)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    total = n * d
    m = total // k
    perm = torch.rand((n_inst, total), generator=g, device=device).argsort(dim=1)[:, : m * k]
    tuples = (perm // d).to(torch.int32).reshape(n_inst, m, k)
    del perm
    srt, _ = torch.sort(tuples, dim=2)
    ok = (srt[:, :, 1:] != srt[:, :, :-1]).all(dim=2)              # drop clauses with a repeated variable
    del srt
    signs = torch.randint(0, 2, tuples.shape, generator=g, device=device, dtype=torch.int32)
    lits = (tuples * 2 + signs)[ok].contiguous()                   # boolean mask keeps instance order
    counts = ok.sum(dim=1).cpu()
    off = torch.zeros(n_inst + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(counts, 0)
    return off, lits


def uniform_ksat_torch(n: int, k: int, m: int, seed: int, device="cuda"):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    vars_ = torch.randint(0, n, (m, k), generator=g, device=device, dtype=torch.int32)
    while True:
        srt, _ = torch.sort(vars_, dim=1)
        bad = (srt[:, 1:] == srt[:, :-1]).any(dim=1)
        nb = int(bad.sum())
        if nb == 0:
            break
        vars_[bad] = torch.randint(0, n, (nb, k), generator=g, device=device, dtype=torch.int32)
    signs = torch.randint(0, 2, (m, k), generator=g, device=device, dtype=torch.int32)
    return (vars_ * 2 + signs).contiguous()


# ---------------------------------------------------------------------------
# DIMACS text helpers (small instances only; cfg1 goes through a DIMACS file)
# ---------------------------------------------------------------------------

def lits_to_signed(lits: np.ndarray) -> np.ndarray:
    """Inverse of example/main.cpp:168: even 2v -> v+1, odd 2v+1 -> -(v+1)."""
    v = (lits >> 1).astype(np.int64) + 1
    return np.where(lits & 1, -v, v).astype(np.int32)


def signed_to_lits(signed: np.ndarray) -> np.ndarray:
    """example/main.cpp:168: x>0 -> 2x-2, else -2x-1."""
    s = signed.astype(np.int64)
    return np.where(s > 0, 2 * s - 2, -2 * s - 1).astype(np.uint32)


def write_dimacs(path: str, n_vars: int, clauses, comment: str | None = None, trailing_newline: bool = True) -> None:
    """``clauses``: (m,k) literal matrix or list of literal lists (encoded, not signed)."""
    lines = []
    if comment:
        lines.append(f"c {comment}")
    lines.append(f"p cnf {n_vars} {len(clauses)}")
    for c in clauses:
        sg = lits_to_signed(np.asarray(c, dtype=np.uint32))
        lines.append(" ".join(str(int(x)) for x in sg) + " 0")
    text = "\n".join(lines) + ("\n" if trailing_newline else "")
    with open(path, "w") as f:
        f.write(text)
