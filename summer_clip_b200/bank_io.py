"""Sidecar files for the kernel-layout key bank.

The reference re-loads `[D, N]` `.pt` feature banks (clip_adapter/save_features.py:28-64) and the `[N, C]` logits
bank (clip_searcher/save_image_outs.py:21-27) on every run and re-normalises them per beta.  Once attention
costs ~0.1 s, loading + normalising + sorting dominate a run, so the normalised, label-sorted, K-major bank the
kernel consumes can be written next to its sources and memory-mapped back:

    <dir>/rows.npy   uint16 [n_rows, D_pad]   fp16 / bf16 bit patterns of HardBank.rows
    <dir>/perm.npy   int64  [n_rows]          sorted position -> original key (-1 = padding)
    <dir>/gcls.npy   int16  [n_rows / 16]
    <dir>/kbits.npy  int32  [n_rows / 32]
    <dir>/meta.json  format version, dtype, n_sorted / n_keys / n_classes, the key it was built for

`bank_key` fingerprints what the bank was built from (source files by path + size + mtime, the selected indices,
class count, operand dtype), so a stale sidecar is never picked up.  Plain numpy I/O: no kernels here.
"""
from __future__ import annotations

import hashlib
import json
import os
import typing as tp
from pathlib import Path

import numpy as np
import torch

from . import ops

FORMAT_VERSION = 1
_DTYPES = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float8_e4m3fn": torch.float8_e4m3fn}


def bank_key(sources: tp.Sequence[tp.Union[str, os.PathLike]], n_classes: int, op_dtype: torch.dtype,
             idx: tp.Optional[torch.Tensor] = None, extra: str = "") -> str:
    """sha256 over the source files' identity (absolute path, size, mtime), the selection and the layout knobs."""
    h = hashlib.sha256()
    h.update(f"v{FORMAT_VERSION}|{n_classes}|{op_dtype}|{extra}".encode())
    for src in sources:
        p = Path(src).resolve()
        st = p.stat()
        h.update(f"|{p}|{st.st_size}|{st.st_mtime_ns}".encode())
    if idx is not None:
        h.update(np.ascontiguousarray(idx.detach().cpu().numpy().astype(np.int64)).tobytes())
    return h.hexdigest()


def save_hard_bank(bank: "ops.HardBank", directory: tp.Union[str, os.PathLike], key: str = "") -> Path:
    """Write a gathered HardBank (rows present) as a sidecar directory; atomic via a temporary name."""
    assert bank.rows is not None, "HardBank.gather() first"
    directory = Path(directory)
    tmp = directory.with_name(directory.name + f".tmp{os.getpid()}")
    tmp.mkdir(parents=True, exist_ok=True)
    rows = bank.rows.detach().contiguous().cpu()
    np.save(tmp / "rows.npy", rows.view(torch.int16).numpy().view(np.uint16))
    np.save(tmp / "perm.npy", bank.perm.detach().cpu().numpy())
    np.save(tmp / "gcls.npy", bank.gcls.detach().cpu().numpy())
    np.save(tmp / "kbits.npy", bank.kbits.detach().cpu().numpy())
    meta = {"format": FORMAT_VERSION, "dtype": str(rows.dtype).replace("torch.", ""), "n_sorted": bank.n_sorted,
            "n_keys": bank.n_keys, "n_classes": bank.n_classes, "key": key}
    (tmp / "meta.json").write_text(json.dumps(meta))
    if directory.exists():
        for f in directory.iterdir():
            f.unlink()
        directory.rmdir()
    tmp.rename(directory)
    return directory


def load_hard_bank(directory: tp.Union[str, os.PathLike], device: tp.Union[str, torch.device] = "cuda",
                   key: tp.Optional[str] = None) -> tp.Optional["ops.HardBank"]:
    """Memory-map a sidecar back.  Returns None when it is absent, of another format version or built for a
    different `key` (the caller then rebuilds and saves)."""
    directory = Path(directory)
    meta_path = directory / "meta.json"
    if not meta_path.exists():
        return None
    meta = json.loads(meta_path.read_text())
    if meta.get("format") != FORMAT_VERSION or (key is not None and meta.get("key") != key) or meta.get("dtype") not in _DTYPES:
        return None
    dev = torch.device(device)

    def put(name: str) -> torch.Tensor:
        arr = np.load(directory / name, mmap_mode="r")          # pages are read once, straight into the copy below
        t = torch.from_numpy(np.array(arr)) if dev.type == "cpu" else torch.from_numpy(np.asarray(arr).copy())
        return t.to(dev, non_blocking=True)

    rows = put("rows.npy").view(torch.int16).view(_DTYPES[meta["dtype"]])
    bank = ops.HardBank(put("perm.npy"), put("gcls.npy"), put("kbits.npy"), meta["n_sorted"], meta["n_keys"], meta["n_classes"])
    bank.rows = rows
    return bank
