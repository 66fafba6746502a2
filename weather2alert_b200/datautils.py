"""Climate-similarity groups, precomputed once into integer tables.

Restates `get_similar_counties` (`/root/reference/src/weather2alert/datautils.py:103-126`):
counties are grouped by their Building-America climate zone, except that every county
of a "western" state collapses into ``"Cold-West"`` whatever its zone
(`datautils.py:113-120`), and the remaining ``"Cold"`` counties become ``"Cold-East"``.
The reference rebuilds that relabelled frame with a row-wise ``.apply`` on every
augmented reset (19 ms); here it is done once, and the result is
a CSR table that the reset path indexes with integers.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np

# State FIPS prefix -> USPS code.  Same facts as the reference's FIPS2STATE table
# (`datautils.py:42-100`), *including* its non-standard "03" -> AZ entry, because the
# grouping below has to agree with the reference county for county.
_STATE_PAIRS = (
    "01AL 02AK 03AZ 04AZ 05AR 06CA 08CO 09CT 10DE 11DC 12FL 13GA 15HI 16ID 17IL 18IN "
    "19IA 20KS 21KY 22LA 23ME 24MD 25MA 26MI 27MN 28MS 29MO 30MT 31NE 32NV 33NH 34NJ "
    "35NM 36NY 37NC 38ND 39OH 40OK 41OR 42PA 44RI 45SC 46SD 47TN 48TX 49UT 50VT 51VA "
    "53WA 54WV 55WI 56WY 72PR 60AS 66GU 69MP 78VI"
)
FIPS2STATE = {p[:2]: p[2:] for p in _STATE_PAIRS.split()}

# `datautils.py:3-17` -- ND, SD, NE, KS are deliberately part of the western group.
WESTERN_STATES = frozenset("AZ CA CO ID MT NM NV OR WA ND SD NE KS".split())


def relabel_zone(fips: str, ba_zone: str) -> str:
    """Zone label used for grouping (`datautils.py:109-120`)."""
    state = FIPS2STATE.get(str(fips)[:2])
    if state in WESTERN_STATES:
        return "Cold-West"
    if ba_zone == "Cold":
        return "Cold-East"
    return ba_zone


class ClimateGroups:
    """CSR table ``county -> similar counties`` in confounders row order.

    Parameters
    ----------
    conf_fips, conf_zone
        the ``fips`` and ``ba_zone`` columns of the confounders frame, in row order
        (the order matters: it is the order of the list the reference samples from,
        `datautils.py:124`).
    fips_list
        the weights' county list; members not in it are dropped, as at
        `/root/reference/src/weather2alert/env.py:116`.
    """

    def __init__(
        self,
        conf_fips: Sequence[str],
        conf_zone: Sequence[str],
        fips_list: Iterable[str],
    ):
        self.conf_fips = [str(f) for f in conf_fips]
        labels = [relabel_zone(f, z) for f, z in zip(self.conf_fips, conf_zone)]
        allowed = set(str(f) for f in fips_list)
        names: dict[str, int] = {}
        members: list[list[str]] = []
        for f, lab in zip(self.conf_fips, labels):
            g = names.setdefault(lab, len(names))
            if g == len(members):
                members.append([])
            if f in allowed:
                members[g].append(f)
        self.group_names = list(names)
        self.group_of = {f: names[lab] for f, lab in zip(self.conf_fips, labels)}
        self.members = members
        self.offsets = np.zeros(len(members) + 1, dtype=np.int64)
        np.cumsum([len(m) for m in members], out=self.offsets[1:])

    def similar(self, fips: str) -> list[str]:
        """Filtered similar-county list of `fips` (`env.py:115-116`).

        Raises ``KeyError`` for a county that is not in the confounders frame, as
        ``confounders.loc[fips]`` does at `datautils.py:123`.
        """
        try:
            return self.members[self.group_of[str(fips)]]
        except KeyError:
            raise KeyError(fips) from None

    def count(self, fips: str) -> int:
        return len(self.similar(fips))

    def counts_for(self, fips_seq: Sequence[str]) -> np.ndarray:
        """``n_c`` for each county of `fips_seq` (0 when absent from confounders)."""
        out = np.zeros(len(fips_seq), dtype=np.int32)
        for i, f in enumerate(fips_seq):
            g = self.group_of.get(str(f))
            out[i] = 0 if g is None else len(self.members[g])
        return out


def get_similar_counties(fips: str, conf_fips: Sequence[str], conf_zone: Sequence[str]) -> list[str]:
    """Unfiltered similar-county list, the exact output of the reference helper."""
    target = None
    labels = []
    for f, z in zip(conf_fips, conf_zone):
        lab = relabel_zone(f, z)
        labels.append(lab)
        if str(f) == str(fips):
            target = lab
    if target is None:
        raise KeyError(fips)
    return [str(f) for f, lab in zip(conf_fips, labels) if lab == target]
