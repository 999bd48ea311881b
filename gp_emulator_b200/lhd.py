"""Latin-hypercube sample designs over ``scipy.stats`` distributions (host utility).

Same call and same results as the reference's ``lhd`` (gp_emulator/lhd.py:11-269, exported from
gp_emulator/__init__.py:3): training-set designs for an emulator are drawn with it.  The legacy global numpy RNG is
consumed in the reference's order -- one uniform draw per (variable, stratum), then one ``randint`` per (column, row)
swap -- so a seeded call reproduces the reference's design exactly (tests/test_utilities.py holds that against frozen
reference outputs).
"""
from __future__ import annotations

import numpy as np


def _stratified_uniform(size, nvars):
    """One point per stratum [i/size, (i+1)/size) and variable, strata in order (reference lhd.py:143-161)."""
    seg = 1.0 / size
    out = np.empty((size, nvars))
    for n in range(nvars):
        out[:, n] = np.arange(size) * seg + np.random.random(size) * seg
    return out


def _shuffle_columns(data):
    """The reference's in-column shuffle: row i swaps with a uniformly drawn row j, i ascending (lhd.py:166-181)."""
    out = data.copy()
    rows, cols = out.shape
    for k in range(cols):
        partner = np.random.randint(rows, size=rows)
        col = out[:, k]
        for i in range(rows):
            j = partner[i]
            col[i], col[j] = col[j], col[i]
    return out


def _inverse_square_energy(points):
    """sum over pairs of 1 / |p_i - p_j|^2 (reference lhd.py:201-208)."""
    diff = points[:, None, :] - points[None, :, :]
    d2 = np.sum(diff * diff, axis=2)
    iu = np.triu_indices(points.shape[0], k=1)
    return float(np.sum(1.0 / d2[iu]))


def lhd(dist=None, size=None, dims=1, form="randomized", iterations=100, showcorrelations=False):
    """Latin-hypercube design, ``size`` rows; one column per distribution in ``dist`` (or ``dims`` columns of one).

    ``dist``: a frozen ``scipy.stats`` distribution or a sequence of them (anything with ``ppf``).
    ``form``: ``'randomized'`` or ``'spacefilling'``.  The space-filling form evaluates the pair energy of
    ``iterations`` successive shuffles; like the reference (whose running best is never updated, lhd.py:211-218) it
    returns the LAST design evaluated, not the minimum-energy one -- kept so seeded designs stay identical.
    ``'orthogonal'`` raises NotImplementedError as in the reference (:244-245).  Returns None without ``dist``/``size``.
    """
    assert dims > 0, 'kwarg "dims" must be at least 1'
    if not size or not dist:
        return None
    if form not in ("randomized", "spacefilling", "orthogonal"):
        raise ValueError('Invalid "form" value: %s' % (form,))
    if form == "orthogonal":
        raise NotImplementedError("Sorry. The orthogonal space-filling algorithm hasn't been implemented yet.")
    dists = list(dist) if hasattr(dist, "__getitem__") else [dist] * dims
    nvars = len(dists)
    unif = _shuffle_columns(_stratified_uniform(size, nvars))
    if form == "spacefilling":
        energy, chosen = None, unif
        for _ in range(iterations):
            energy, chosen = _inverse_square_energy(unif), unif
            unif = _shuffle_columns(unif)
        if iterations > 0:
            print("Optimized Distance:", energy)
        unif = chosen
    design = np.empty_like(unif)
    for i, d in enumerate(dists):
        design[:, i] = d.ppf(unif[:, i])
    if showcorrelations and nvars > 1:
        cor = np.corrcoef(design, rowvar=False)
        inv = np.linalg.pinv(cor)
        print("Correlation Matrix:\n", cor)
        print("Inverted Correlation Matrix:\n", inv)
        print("Variance Inflation Factor (VIF):", np.max(np.diag(inv)))
    return design
