"""Sidecar files for the kernel-layout key bank.

The reference re-loads `[D, N]` `.pt` feature banks (clip_adapter/save_features.py:28-64) and the `[N, C]` logits
bank (clip_searcher/save_image_outs.py:21-27) on every run and re-normalises them per beta.  Once attention
costs ~0.1 s, loading + normalising + sorting dominate a run, so the normalised, label-sorted, K-major bank the
kernel consumes can be written next to its sources and memory-mapped back:

    <dir>/rows.npy   uint16 [n_rows, D_pad]   fp16 / bf16 bit patterns of HardBank.rows
    <dir>/perm.npy   int64  [n_rows]          sorted position -> original key (-1 = padding)
    <dir>/gcls.npy   int16  [n_rows / 16]
    <dir>/kbits.npy  int32  [n_rows / 32]
    <dir>/meta.json  format version, kind, dtype, n_sorted / n_keys / n_classes, the key it was built for

Dense-value caches (SoftmaxCacheStrategy) keep `rows.npy` = the normalised keys in their original order plus
`vt.npy` (uint16 [C_pad, Nk_pad], the transposed values); query banks (`save_query_bank`) keep the normalised query
rows and, optionally, the zero-shot logits `z.npy` — the reference recomputes both for every run of a sweep
(image_attention.py:72-83).

Loading never builds a second host copy of a bank: the `.npy` payload is memory-mapped and streamed to the device
in 64 MB pieces through two pinned staging buffers, the copy of piece i overlapping the page-in of piece i + 1.

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

FORMAT_VERSION = 2
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
    meta = {"format": FORMAT_VERSION, "kind": "hard", "dtype": str(bank.rows.dtype).replace("torch.", ""),
            "n_sorted": bank.n_sorted, "n_keys": bank.n_keys, "n_classes": bank.n_classes, "key": key}
    arrays = {"rows.npy": _rows_to_npy(bank.rows), "perm.npy": bank.perm.detach().cpu().numpy(),
              "gcls.npy": bank.gcls.detach().cpu().numpy(), "kbits.npy": bank.kbits.detach().cpu().numpy()}
    return _write_dir(Path(directory), arrays, meta)


_STAGE_BYTES = 64 << 20


def _to_device(path: Path, device: torch.device) -> torch.Tensor:
    """One `.npy` array onto `device`.  CUDA: memory-map the file and stream it through two pinned staging buffers
    (no full host copy; the H2D of a piece overlaps the page-in of the next).  CPU: one copy out of the map."""
    arr = np.load(path, mmap_mode="r")
    if device.type == "cpu":
        host = np.array(arr)
        return torch.from_numpy(host.view(np.int16) if host.dtype == np.uint16 else host)
    flat = arr.reshape(-1).view(np.uint8) if arr.dtype != np.uint8 else arr.reshape(-1)
    out = torch.empty(flat.shape[0], dtype=torch.uint8, device=device)
    n = flat.shape[0]
    if n:
        stage = [torch.empty(min(_STAGE_BYTES, n), dtype=torch.uint8).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        with torch.cuda.device(device):
            for i, s0 in enumerate(range(0, n, _STAGE_BYTES)):
                s1 = min(n, s0 + _STAGE_BYTES)
                buf = stage[i % 2]
                if i >= 2:
                    done[i % 2].synchronize()              # the copy that last read this staging buffer has finished
                buf[: s1 - s0].numpy()[:] = flat[s0:s1]    # page-in + memcpy into pinned memory
                out[s0:s1].copy_(buf[: s1 - s0], non_blocking=True)
                done[i % 2].record()
            torch.cuda.current_stream().synchronize()
    np_dtype = arr.dtype
    tdtype = {np.dtype(np.uint16): torch.int16, np.dtype(np.int16): torch.int16, np.dtype(np.int32): torch.int32,
              np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32, np.dtype(np.uint8): torch.uint8}[np_dtype]
    return out.view(tdtype).view(tuple(arr.shape))


def _rows_to_npy(t: torch.Tensor) -> np.ndarray:
    t = t.detach().contiguous().cpu()
    if t.element_size() == 1:
        return t.view(torch.uint8).numpy()
    return t.view(torch.int16).numpy().view(np.uint16)


def _rows_from(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return t.view(torch.uint8).view(dtype) if dtype.itemsize == 1 else t.view(torch.int16).view(dtype)


def _write_dir(directory: Path, arrays: tp.Mapping[str, np.ndarray], meta: dict) -> Path:
    """Atomic directory write: everything under a temporary name, then one rename."""
    directory = Path(directory)
    tmp = directory.with_name(directory.name + f".tmp{os.getpid()}")
    tmp.mkdir(parents=True, exist_ok=True)
    for name, arr in arrays.items():
        np.save(tmp / name, arr)
    (tmp / "meta.json").write_text(json.dumps(meta))
    if directory.exists():
        for f in directory.iterdir():
            f.unlink()
        directory.rmdir()
    tmp.rename(directory)
    return directory


def _read_meta(directory: Path, kind: str, key: tp.Optional[str]) -> tp.Optional[dict]:
    meta_path = Path(directory) / "meta.json"
    if not meta_path.exists():
        return None
    meta = json.loads(meta_path.read_text())
    if meta.get("format") != FORMAT_VERSION or meta.get("kind") != kind or meta.get("dtype") not in _DTYPES:
        return None
    if key is not None and meta.get("key") != key:
        return None
    return meta


def save_dense_bank(k_norm: torch.Tensor, vt: torch.Tensor, n_keys: int, n_classes: int,
                    directory: tp.Union[str, os.PathLike], key: str = "") -> Path:
    """Dense-value cache: normalised keys [n_keys, D_pad] + transposed values Vt [C_pad, Nk_pad] (both op dtype)."""
    meta = {"format": FORMAT_VERSION, "kind": "dense", "dtype": str(k_norm.dtype).replace("torch.", ""), "n_keys": int(n_keys),
            "n_classes": int(n_classes), "key": key}
    return _write_dir(Path(directory), {"rows.npy": _rows_to_npy(k_norm), "vt.npy": _rows_to_npy(vt)}, meta)


def load_dense_bank(directory: tp.Union[str, os.PathLike], device: tp.Union[str, torch.device] = "cuda",
                    key: tp.Optional[str] = None):
    """(k_norm, vt, n_keys, n_classes) or None."""
    directory = Path(directory)
    meta = _read_meta(directory, "dense", key)
    if meta is None:
        return None
    dev, dt = torch.device(device), _DTYPES[meta["dtype"]]
    return (_rows_from(_to_device(directory / "rows.npy", dev), dt), _rows_from(_to_device(directory / "vt.npy", dev), dt),
            meta["n_keys"], meta["n_classes"])


def save_query_bank(q_norm: torch.Tensor, directory: tp.Union[str, os.PathLike], key: str = "",
                    clip_logits: tp.Optional[torch.Tensor] = None) -> Path:
    """Normalised query rows [Nq, D_pad] (op dtype) and, optionally, their zero-shot logits Z [Nq, C] fp32."""
    arrays = {"rows.npy": _rows_to_npy(q_norm)}
    if clip_logits is not None:
        arrays["z.npy"] = clip_logits.detach().float().contiguous().cpu().numpy()
    meta = {"format": FORMAT_VERSION, "kind": "queries", "dtype": str(q_norm.dtype).replace("torch.", ""),
            "n_queries": int(q_norm.shape[0]), "has_logits": clip_logits is not None, "key": key}
    return _write_dir(Path(directory), arrays, meta)


def load_query_bank(directory: tp.Union[str, os.PathLike], device: tp.Union[str, torch.device] = "cuda",
                    key: tp.Optional[str] = None):
    """(q_norm, clip_logits or None) or None."""
    directory = Path(directory)
    meta = _read_meta(directory, "queries", key)
    if meta is None:
        return None
    dev = torch.device(device)
    q = _rows_from(_to_device(directory / "rows.npy", dev), _DTYPES[meta["dtype"]])
    z = _to_device(directory / "z.npy", dev) if meta.get("has_logits") else None
    return q, z


def load_hard_bank(directory: tp.Union[str, os.PathLike], device: tp.Union[str, torch.device] = "cuda",
                   key: tp.Optional[str] = None) -> tp.Optional["ops.HardBank"]:
    """Stream a label-sorted bank sidecar back onto `device`.  Returns None when it is absent, of another format
    version / kind or built for a different `key` (the caller then rebuilds and saves)."""
    directory = Path(directory)
    meta = _read_meta(directory, "hard", key)
    if meta is None:
        return None
    dev = torch.device(device)
    bank = ops.HardBank(_to_device(directory / "perm.npy", dev), _to_device(directory / "gcls.npy", dev),
                        _to_device(directory / "kbits.npy", dev), meta["n_sorted"], meta["n_keys"], meta["n_classes"])
    bank.rows = _rows_from(_to_device(directory / "rows.npy", dev), _DTYPES[meta["dtype"]])
    return bank
