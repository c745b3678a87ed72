"""Kubernetes resource.Quantity parsing (the subset a scheduler needs): decimal/binary SI suffixes, milli."""
from __future__ import annotations

from decimal import Decimal, InvalidOperation

_BIN = {"Ki": 1024, "Mi": 1024 ** 2, "Gi": 1024 ** 3, "Ti": 1024 ** 4, "Pi": 1024 ** 5, "Ei": 1024 ** 6}
_DEC = {"n": Decimal("1e-9"), "u": Decimal("1e-6"), "m": Decimal("1e-3"), "": Decimal(1), "k": Decimal("1e3"), "M": Decimal("1e6"), "G": Decimal("1e9"),
        "T": Decimal("1e12"), "P": Decimal("1e15"), "E": Decimal("1e18")}


def parse_quantity(q) -> Decimal:
    if isinstance(q, (int, float, Decimal)):
        return Decimal(q)
    s = str(q).strip()
    if not s:
        raise ValueError("empty quantity")
    for suf, mult in _BIN.items():
        if s.endswith(suf):
            return Decimal(s[:-2]) * mult
    suffix = s[-1] if s[-1] in _DEC and not s[-1].isdigit() else ""
    num = s[:-1] if suffix else s
    try:
        return Decimal(num) * _DEC[suffix]
    except InvalidOperation as e:
        raise ValueError(f"invalid quantity {q!r}") from e
