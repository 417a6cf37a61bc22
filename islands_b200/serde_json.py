"""serde_json forms of the reference's plain-data types — the `#[derive(Serialize, Deserialize)]` of `DistanceMetric`
(distance.rs:7-19, `rename_all = "lowercase"`), `PruningStrategy` (leann.rs:166-177), `LeannConfig` (leann.rs:321-375),
`HnswConfig` (hnsw.rs:13-28), `PQConfig` (pq.rs:12-22), `SearchConfig` (search.rs:7-20) and `SearchResult`
(search.rs:54-67) — so that a configuration or result written by the reference reads back here and the other way round.
Host-side format code only.

serde_json writes a struct as an object with the fields in declaration order and no whitespace, a unit enum variant as
its (renamed) name, `Option::None` as `null`, integers in decimal, and floats with the shortest digits that round-trip
(the `ryu` crate) — `f32` fields with the shortest digits for the *f32* value, which is why `0.02f32` is written `0.02`
and not `0.019999999552965164`.  `ryu`'s layout rule (decimal notation while the decimal exponent is in (-5, 16] for
f64 / (-6, 13] for f32, exponent notation `1.234e33` / `1e-7` outside, always at least one fractional digit) is restated
in `format_float`; it is third-party to the reference and pinned by restatement only.  Non-finite floats are `null`,
as serde_json writes them.
"""
import json
import math

import numpy as np

from .core import CoreError, DistanceMetric, HnswConfig, LeannConfig, PQConfig, PruningStrategy
from .search import SearchConfig, SearchResult
from .storage import DeserializationError  # CoreError::Deserialization: malformed JSON, unknown variant, bad field


METRIC_NAMES = {DistanceMetric.Cosine: "cosine", DistanceMetric.Euclidean: "euclidean",
                DistanceMetric.DotProduct: "dotproduct", DistanceMetric.Manhattan: "manhattan"}
STRATEGY_NAMES = {PruningStrategy.Global: "Global", PruningStrategy.Local: "Local",
                  PruningStrategy.Proportional: "Proportional"}


def format_float(x, f32=False) -> str:
    """One float as serde_json (ryu) prints it; `f32` picks the shortest digits of the value as a 32-bit float."""
    v = np.float32(x) if f32 else np.float64(x)
    if not np.isfinite(v):
        return "null"
    if v == 0:
        return "-0.0" if math.copysign(1.0, float(v)) < 0 else "0.0"
    sci = np.format_float_scientific(v, unique=True, trim="-")  # e.g. '-1.2345e+10', '3e-01'
    mant, exp = sci.split("e")
    sign = "-" if mant.startswith("-") else ""
    digits = mant.lstrip("-").replace(".", "")
    length, kk = len(digits), int(exp) + 1      # 10^(kk-1) <= |v| < 10^kk
    k = kk - length                             # |v| = digits * 10^k
    hi, lo = (13, -6) if f32 else (16, -5)
    if 0 <= k and kk <= hi:                     # 1234e7 -> 12340000000.0
        body = digits + "0" * k + ".0"
    elif 0 < kk <= hi:                          # 1234e-2 -> 12.34
        body = digits[:kk] + "." + digits[kk:]
    elif lo < kk <= 0:                          # 1234e-6 -> 0.001234
        body = "0." + "0" * (-kk) + digits
    elif length == 1:                           # 1e30
        body = f"{digits}e{kk - 1}"
    else:                                       # 1234e30 -> 1.234e33
        body = f"{digits[0]}.{digits[1:]}e{kk - 1}"
    return sign + body


def _obj(pairs) -> str:
    return "{" + ",".join(json.dumps(k) + ":" + v for k, v in pairs) + "}"


def _bool(b) -> str:
    return "true" if b else "false"


def _value(x):
    return int(getattr(x, "value", x))


# ---- to_json ----------------------------------------------------------------------------------------------------------

def metric_to_json(metric) -> str:
    return json.dumps(METRIC_NAMES[_value(metric)])


def strategy_to_json(strategy) -> str:
    return json.dumps(STRATEGY_NAMES[_value(strategy)])


def leann_config_to_json(c: LeannConfig) -> str:
    pairs = [("m", str(c.m)), ("m0", str(c.m0)), ("ef_construction", str(c.ef_construction)), ("ml", format_float(c.ml)),
             ("max_layers", str(c.max_layers)), ("metric", metric_to_json(c.metric)), ("ef_search", str(c.ef_search)),
             ("beam_width", str(c.beam_width)), ("prune_ratio", format_float(c.prune_ratio, f32=True)),
             ("pruning_strategy", strategy_to_json(c.pruning_strategy)),
             ("high_degree_pruning", _bool(c.high_degree_pruning)),
             ("hub_percentile", format_float(c.hub_percentile, f32=True)), ("is_compact", _bool(c.is_compact)),
             ("is_recompute", _bool(c.is_recompute))]
    if c.prune_seed:  # not a reference field (include/islands_b200.h); serde ignores unknown keys when reading
        pairs.append(("prune_seed", str(c.prune_seed)))
    return _obj(pairs)


def hnsw_config_to_json(c: HnswConfig) -> str:
    return _obj([("m", str(c.m)), ("m0", str(c.m0)), ("ef_construction", str(c.ef_construction)),
                 ("ml", format_float(c.ml)), ("metric", metric_to_json(c.metric)), ("max_layers", str(c.max_layers))])


def pq_config_to_json(c: PQConfig) -> str:
    return _obj([("num_subquantizers", str(c.num_subquantizers)), ("num_centroids", str(c.num_centroids)),
                 ("training_iterations", str(c.training_iterations)),
                 ("seed", "null" if c.seed is None else str(c.seed))])


def search_config_to_json(c: SearchConfig) -> str:
    return _obj([("top_k", str(int(c.top_k))), ("ef", str(int(c.ef))), ("include_vectors", _bool(c.include_vectors)),
                 ("include_metadata", _bool(c.include_metadata)),
                 ("min_similarity", "null" if c.min_similarity is None else format_float(c.min_similarity, f32=True))])


def search_result_to_json(r: SearchResult) -> str:
    vector = "null" if r.vector is None else "[" + ",".join(format_float(x, f32=True) for x in np.asarray(r.vector).ravel()) + "]"
    metadata = "null" if r.metadata is None else json.dumps(r.metadata, separators=(",", ":"), ensure_ascii=False)
    text = "null" if r.text is None else json.dumps(r.text, ensure_ascii=False)
    return _obj([("id", str(int(r.id))), ("score", format_float(r.score, f32=True)), ("vector", vector),
                 ("metadata", metadata), ("text", text)])


# ---- from_json --------------------------------------------------------------------------------------------------------

def _load(text):
    try:
        return json.loads(text)
    except (ValueError, TypeError) as e:
        raise DeserializationError(str(e)) from None


def _field(o, name, kind, optional=False):
    if not isinstance(o, dict):
        raise DeserializationError("expected a JSON object")
    if name not in o:
        if optional:  # Option<T>: a missing key deserialises as None
            return None
        raise DeserializationError(f"missing field `{name}`")
    v = o[name]
    if v is None and optional:
        return None
    if kind == "uint":
        if isinstance(v, bool) or not isinstance(v, int) or v < 0:
            raise DeserializationError(f"field `{name}`: expected an unsigned integer")
    elif kind == "float":
        if isinstance(v, bool) or not isinstance(v, (int, float)):
            raise DeserializationError(f"field `{name}`: expected a number")
        v = float(v)
    elif kind == "bool":
        if not isinstance(v, bool):
            raise DeserializationError(f"field `{name}`: expected a boolean")
    elif kind == "str":
        if not isinstance(v, str):
            raise DeserializationError(f"field `{name}`: expected a string")
    return v


def _variant(v, names, what):
    for code, name in names.items():
        if v == name:
            return code
    raise DeserializationError(f"unknown variant `{v}` of {what}, expected one of " + ", ".join(f"`{n}`" for n in names.values()))


def metric_from_json(text) -> DistanceMetric:
    return DistanceMetric(_variant(_load(text), METRIC_NAMES, "DistanceMetric"))


def strategy_from_json(text) -> int:
    return _variant(_load(text), STRATEGY_NAMES, "PruningStrategy")


def leann_config_from_json(text) -> LeannConfig:
    o = _load(text)
    c = LeannConfig()
    for name in ("m", "m0", "ef_construction", "max_layers", "ef_search", "beam_width"):
        setattr(c, name, _field(o, name, "uint"))
    for name in ("ml", "prune_ratio", "hub_percentile"):
        setattr(c, name, _field(o, name, "float"))
    for name in ("high_degree_pruning", "is_compact", "is_recompute"):
        setattr(c, name, int(_field(o, name, "bool")))
    c.metric = _variant(_field(o, "metric", "str"), METRIC_NAMES, "DistanceMetric")
    c.pruning_strategy = _variant(_field(o, "pruning_strategy", "str"), STRATEGY_NAMES, "PruningStrategy")
    seed = _field(o, "prune_seed", "uint", optional=True)
    c.prune_seed = 0 if seed is None else seed
    return c


def hnsw_config_from_json(text) -> HnswConfig:
    o = _load(text)
    return HnswConfig(m=_field(o, "m", "uint"), m0=_field(o, "m0", "uint"),
                      ef_construction=_field(o, "ef_construction", "uint"), ml=_field(o, "ml", "float"),
                      metric=_variant(_field(o, "metric", "str"), METRIC_NAMES, "DistanceMetric"),
                      max_layers=_field(o, "max_layers", "uint"))


def pq_config_from_json(text) -> PQConfig:
    o = _load(text)
    return PQConfig(_field(o, "num_subquantizers", "uint"), _field(o, "num_centroids", "uint"),
                    _field(o, "training_iterations", "uint"), _field(o, "seed", "uint", optional=True))


def search_config_from_json(text) -> SearchConfig:
    o = _load(text)
    ms = _field(o, "min_similarity", "float", optional=True)
    return SearchConfig(_field(o, "top_k", "uint"), _field(o, "ef", "uint"), _field(o, "include_vectors", "bool"),
                        _field(o, "include_metadata", "bool"), None if ms is None else np.float32(ms))


def search_result_from_json(text) -> SearchResult:
    o = _load(text)
    vec = o.get("vector") if isinstance(o, dict) else None
    if vec is not None:
        if not isinstance(vec, list) or any(isinstance(x, bool) or not isinstance(x, (int, float)) for x in vec):
            raise DeserializationError("field `vector`: expected an array of numbers")
        vec = np.asarray(vec, np.float32)
    return SearchResult(_field(o, "id", "uint"), _field(o, "score", "float"), vec, o.get("metadata"),
                        _field(o, "text", "str", optional=True))


__all__ = [n for n in dir() if n.endswith("_json") or n in ("format_float", "DeserializationError")]
assert issubclass(DeserializationError, CoreError)
