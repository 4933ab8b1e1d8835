"""Host-side helpers with the reference's names (videoprism/utils.py): checkpoint trees and text canonicalisation.

Pure Python / numpy: none of this is on the device path.  The tree helpers are what `models.load_pretrained_weights`
(models.py:306-336) uses to turn a `flax_*_repeated.npz` (flat '/'-joined keys) into the nested `{'params': ...}` state
that `model.apply` takes; `vp_set_weight` consumes the same '/'-joined keys.
"""
from __future__ import annotations

import os
import string
from typing import Any, Dict, Iterable, Iterator, List, Mapping, Sequence, Tuple

import numpy as np

_PUNCT_TO_SPACE = str.maketrans({ch: " " for ch in string.punctuation})


def traverse_with_names(tree, with_inner_nodes: bool = False) -> Iterator[Tuple[str, Any]]:
    """utils.py:30-59: depth-first walk over nested dicts / sequences in sorted-key order, yielding ('a/b/c', leaf)."""
    stack: List[Tuple[str, Any, bool]] = [("", tree, False)]
    while stack:
        name, node, emit_inner = stack.pop()
        if node is None:
            continue
        if emit_inner:
            yield name, node
            continue
        is_map = isinstance(node, Mapping)
        is_seq = isinstance(node, Sequence) and not isinstance(node, (str, bytes))
        if not (is_map or is_seq):
            yield name, node
            continue
        if with_inner_nodes:          # inner node comes after all of its children (post-order), as in the reference
            stack.append((name, node, True))
        children = [(k, node[k]) for k in sorted(node.keys())] if is_map else list(enumerate(node))
        for k, child in reversed(children):
            stack.append((f"{name}/{k}" if name else str(k), child, False))


def tree_flatten_with_names(tree) -> List[Tuple[str, Any]]:
    """utils.py:62-81: [(name, leaf)] in jax.tree.flatten order, which for dict trees is the sorted-key order."""
    return list(traverse_with_names(tree))


def recover_tree(keys: Iterable[str], values: Iterable[Any]) -> Dict[str, Any]:
    """utils.py:84-105: flat '/'-separated names + values -> nested dict."""
    tree: Dict[str, Any] = {}
    for key, value in zip(keys, values):
        node = tree
        *parents, leaf = key.split("/")
        for part in parents:
            node = node.setdefault(part, {})
        node[leaf] = value
    return tree


def npload(fname: str):
    """utils.py:145-154 without the remote cache (no network on the target boxes): np.save or np.savez file."""
    if not os.path.exists(fname):
        raise FileNotFoundError(fname)
    loaded = np.load(fname, allow_pickle=False)
    if isinstance(loaded, np.ndarray):
        return loaded
    with loaded:
        return {k: loaded[k] for k in loaded.files}


def load_checkpoint(npz) -> Dict[str, Any]:
    """utils.py:157-169: path to a .npz (or an already loaded dict-like) -> nested tree."""
    if isinstance(npz, (str, os.PathLike)):
        npz = npload(os.fspath(npz))
    return recover_tree(npz.keys(), npz.values())


def canonicalize_text(text: str) -> str:
    """utils.py:172-201: punctuation -> space, lower case, single spaces, trailing period
    ("Hello, World!" -> "hello world.", utils_test.py:23-26)."""
    words = text.translate(_PUNCT_TO_SPACE).lower().split()
    return " ".join(words) + "."
