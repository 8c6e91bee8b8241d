"""JSON-lines result log with the record schema of the reference run directory.

The reference logs through python-json-logger ('%(asctime)s %(name)s %(levelname)s %(message)s',
conf/hydra_setup.yaml:4-11): a dict message becomes top-level keys with "message": null.  Its
notebooks parse exactly those keys (result_tables.ipynb cells 3, 5, 12), so the same objects are
written here.  W&B (log_utils.py:44-49) needs the network and is not mirrored.
"""
from __future__ import annotations

import json
import sys
import time
import typing as tp
from pathlib import Path


class JsonLinesLogger:
    def __init__(self, name: str, path: tp.Optional[tp.Union[str, Path]] = None, echo: bool = False) -> None:
        self.name = name
        self.path = Path(path) if path is not None else None
        self.echo = echo
        self.records: tp.List[dict] = []
        if self.path is not None:
            self.path.parent.mkdir(parents=True, exist_ok=True)
            self._fh = open(self.path, "a")
        else:
            self._fh = None

    def _emit(self, payload: dict) -> None:
        now = time.time()
        rec = {"asctime": time.strftime("%Y-%m-%d %H:%M:%S", time.localtime(now)) + f",{int(now % 1 * 1000):03d}",
               "name": self.name, "levelname": "INFO"}
        rec.update(payload)
        self.records.append(rec)
        line = json.dumps(rec)
        if self._fh is not None:
            self._fh.write(line + "\n")
            self._fh.flush()
        if self.echo:
            print(line, file=sys.stderr)

    def log_info(self, message: tp.Union[str, dict]) -> None:
        if isinstance(message, dict):
            self._emit({"message": None, **message})
        else:
            self._emit({"message": str(message)})

    # the reference mirrors searcher results to W&B as well (image_attention.py:120)
    log_info_wandb = log_info

    def close(self) -> None:
        if self._fh is not None:
            self._fh.close()
            self._fh = None
