"""SigMF captures as input of the search (row f4 of the scope table: ingest formats).

A SigMF recording is a `<name>.sigmf-data` file of raw samples plus a `<name>.sigmf-meta` JSON
object whose `global` section names the sample type (`core:datatype`) and rate (`core:sample_rate`)
and whose first capture segment may carry the center frequency.  The three complex little-endian
types an LTE capture comes in map onto the engine's wire formats without a conversion pass on the
host: cf32_le -> LTB_FMT_FC32, ci16_le -> LTB_FMT_SC16, ci8 -> LTB_FMT_SC8 (the engine scales sc16
by 2^-15 and sc8 by 2^-7 on the device, as a flowgraph's interleaved_short_to_complex would).
"""
import json
import os

import numpy as np

from . import _abi as A

_TYPES = {
    "cf32_le": (A.FMT_FC32, np.dtype("<c8"), 1), "cf32": (A.FMT_FC32, np.dtype("<c8"), 1),
    "ci16_le": (A.FMT_SC16, np.dtype("<i2"), 2), "ci16": (A.FMT_SC16, np.dtype("<i2"), 2),
    "ci8": (A.FMT_SC8, np.dtype("i1"), 2), "ci8_le": (A.FMT_SC8, np.dtype("i1"), 2),
}


class SigMFError(ValueError):
    pass


def paths(name):
    """(meta path, data path) for `name` given with either extension or none."""
    base = name
    for ext in (".sigmf-meta", ".sigmf-data", ".sigmf"):
        if name.endswith(ext):
            base = name[:-len(ext)]
    return base + ".sigmf-meta", base + ".sigmf-data"


def is_sigmf(name):
    return name.endswith((".sigmf-meta", ".sigmf-data")) or os.path.isfile(name + ".sigmf-meta")


def load(name):
    """-> dict(samples=[n] complex64 or [n, 2] int16 / int8 (memory-mapped), input_format, sample_rate,
    frequency or None, datatype, meta).  Multi-channel recordings and big-endian / real / unsigned
    types are refused: the search takes one complex baseband stream."""
    meta_path, data_path = paths(name)
    try:
        with open(meta_path) as f:
            meta = json.load(f)
    except (OSError, ValueError) as e:
        raise SigMFError("cannot read SigMF metadata %s: %s" % (meta_path, e))
    g = meta.get("global", {})
    dt = g.get("core:datatype")
    if dt not in _TYPES:
        raise SigMFError("SigMF datatype %r is not one of %s" % (dt, ", ".join(sorted(set(_TYPES)))))
    if int(g.get("core:num_channels", 1)) != 1:
        raise SigMFError("multi-channel SigMF recordings are not supported (core:num_channels = %s)" % g.get("core:num_channels"))
    if "core:sample_rate" not in g:
        raise SigMFError("SigMF metadata has no core:sample_rate")
    fmt, dtype, per = _TYPES[dt]
    caps = meta.get("captures") or [{}]
    skip = int(caps[0].get("core:header_bytes", 0)) + int(g.get("core:offset", 0)) * 0
    n_items = (os.path.getsize(data_path) - skip) // (dtype.itemsize * per)
    raw = np.memmap(data_path, dtype=dtype, mode="r", offset=skip, shape=(n_items * per,))
    samples = raw if per == 1 else raw.reshape(n_items, 2)
    return {"samples": samples, "input_format": fmt, "sample_rate": float(g["core:sample_rate"]),
            "frequency": caps[0].get("core:frequency"), "datatype": dt, "meta": meta}


def write(name, samples, sample_rate, frequency=None, description="written by ltetrigger_b200.sigmf"):
    """Write `samples` (complex64, or [n, 2] int16 / int8) as a SigMF recording; returns the two paths."""
    samples = np.asarray(samples)
    dt = {np.dtype("complex64"): "cf32_le", np.dtype("int16"): "ci16_le", np.dtype("int8"): "ci8"}.get(samples.dtype)
    if dt is None or (dt != "cf32_le" and (samples.ndim != 2 or samples.shape[1] != 2)):
        raise SigMFError("samples must be complex64 [n], or int16 / int8 [n, 2]")
    meta_path, data_path = paths(name)
    np.ascontiguousarray(samples).tofile(data_path)
    cap = {"core:sample_start": 0}
    if frequency is not None:
        cap["core:frequency"] = float(frequency)
    meta = {"global": {"core:datatype": dt, "core:sample_rate": float(sample_rate), "core:version": "1.0.0",
                       "core:num_channels": 1, "core:description": description},
            "captures": [cap], "annotations": []}
    with open(meta_path, "w") as f:
        json.dump(meta, f, indent=2)
    return meta_path, data_path
