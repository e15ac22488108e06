"""List <-> padded conversions of per-cloud tensors (gradslam/structures/structutils.py:47-124).

The map store of this package is capacity-backed, so `Pointclouds` itself never needs these (its `*_list` / `*_padded`
properties are views of one store); they are kept because callers and the reference's tests use them directly.
Host-side tensor plumbing only - no arithmetic."""
from typing import List, Optional, Sequence, Union

import torch

__all__ = ["list_to_padded", "padded_to_list"]


def list_to_padded(x: List[torch.Tensor], pad_size: Union[Sequence[int], None] = None, pad_value: float = 0.0,
                   equisized: bool = False) -> torch.Tensor:
    """B tensors (N_b, C_b) -> one (B, rows, cols) tensor filled with `pad_value` outside the items.
    rows / cols come from `pad_size`, else from the largest non-empty item; `equisized=True` just stacks."""
    if equisized:
        return torch.stack(list(x), dim=0)
    if pad_size is not None:
        if len(pad_size) != 2:
            raise ValueError("Pad size must contain target size for 1st and 2nd dim")
        rows, cols = int(pad_size[0]), int(pad_size[1])
    else:
        filled = [t for t in x if len(t) > 0]
        rows = max(t.shape[0] for t in filled)
        cols = max(t.shape[1] for t in filled)
    out = x[0].new_full((len(x), rows, cols), pad_value)
    for slot, item in zip(out, x):
        if len(item) == 0:
            continue
        if item.ndim != 2:
            raise ValueError("Supports only 2-dimensional tensor items")
        slot[: item.shape[0], : item.shape[1]] = item
    return out


def padded_to_list(x: torch.Tensor, split_size: Optional[Sequence] = None) -> List[torch.Tensor]:
    """(B, N, C) -> list of B tensors, item b cut to split_size[b] (an int = rows, or a (rows, cols) pair)."""
    if x.ndim != 3:
        raise ValueError("Supports only 3-dimensional input tensors")
    items = list(torch.unbind(x, dim=0))
    if split_size is None:
        return items
    if len(split_size) != x.shape[0]:
        raise ValueError("Split size must be of same length as inputs first dimension")
    cut = []
    for item, size in zip(items, split_size):
        if isinstance(size, int):
            cut.append(item[:size])
        elif len(size) == 2:
            cut.append(item[: size[0], : size[1]])
        else:
            raise ValueError("Support only for 2-dimensional unbinded tensor. Split size for more dimensions provided")
    return cut
