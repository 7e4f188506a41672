"""Minimal stand-in for ``stable_baselines.logger`` (imported at base_callbacks.py:13): key/value buffer with ``logkv`` /
``dumpkvs`` / ``configure`` / ``get_dir``; output goes to stdout (and ``progress.csv`` under the configured directory)."""
from __future__ import annotations

import csv
import os
from collections import OrderedDict

_dir = None
_kvs: "OrderedDict[str, object]" = OrderedDict()
DEBUG, INFO, WARN, ERROR, DISABLED = 10, 20, 30, 40, 50
_level = INFO


def configure(folder=None, format_strs=None):
    global _dir
    _dir = folder
    if folder:
        os.makedirs(folder, exist_ok=True)


def get_dir():
    return _dir


def set_level(level):
    global _level
    _level = level


def logkv(key, val):
    _kvs[key] = val


record_tabular = logkv


def logkvs(d):
    for k, v in d.items():
        logkv(k, v)


def getkvs():
    return _kvs


def dumpkvs():
    if not _kvs:
        return
    if _level <= INFO:
        w = max(len(str(k)) for k in _kvs)
        print("\n".join("| {:<{w}} | {:<12} |".format(str(k), str(v)[:12], w=w) for k, v in _kvs.items()))
    if _dir:
        path = os.path.join(_dir, "progress.csv")
        new = not os.path.exists(path)
        with open(path, "a", newline="") as f:
            wr = csv.DictWriter(f, fieldnames=list(_kvs.keys()))
            if new:
                wr.writeheader()
            wr.writerow(_kvs)
    _kvs.clear()


dump_tabular = dumpkvs


def log(*args, level=INFO):
    if _level <= level:
        print(*args)


def info(*args):
    log(*args, level=INFO)


def warn(*args):
    log(*args, level=WARN)


def error(*args):
    log(*args, level=ERROR)
