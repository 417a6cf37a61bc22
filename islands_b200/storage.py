"""Index persistence container of the reference (src/core/storage.rs): a file is a sequence of chunks
`tag[4] | len u64 LE | data[len]`; the first chunk is `META` holding the serde_json form of
`IndexMetadata`.  Host-side format code only (no device work); the index payloads that follow are
the `to_bytes()` images of LeannIndex / HnswGraph / ProductQuantizer (islands_b200.core).

Field order and spelling follow the struct (storage.rs:15-29) because serde_json writes fields in
declaration order, compact, UTF-8 unescaped: the bytes written here are the bytes the reference
writes for the same metadata.
"""
import io
import json
import os
import struct
import time
from dataclasses import dataclass
from typing import Optional

from .core import SerializationError


class DeserializationError(SerializationError):
    """CoreError::Deserialization (error.rs) — wrong tag, truncated chunk or malformed JSON."""


@dataclass
class IndexMetadata:
    """storage.rs:15-48."""
    version: int
    num_vectors: int
    dimension: int
    created_at: int
    updated_at: int
    description: Optional[str] = None

    CURRENT_VERSION = 1

    @classmethod
    def new(cls, num_vectors: int, dimension: int) -> "IndexMetadata":
        now = int(time.time())  # chrono::Utc::now().timestamp()
        return cls(cls.CURRENT_VERSION, int(num_vectors), int(dimension), now, now, None)

    def to_json(self) -> bytes:
        """serde_json::to_vec: declaration order, no whitespace, non-ASCII left as UTF-8."""
        fields = (("version", self.version), ("num_vectors", self.num_vectors), ("dimension", self.dimension),
                  ("created_at", self.created_at), ("updated_at", self.updated_at), ("description", self.description))
        return ("{" + ",".join(json.dumps(k) + ":" + json.dumps(v, ensure_ascii=False) for k, v in fields) + "}").encode("utf-8")

    @classmethod
    def from_json(cls, data: bytes) -> "IndexMetadata":
        try:
            o = json.loads(bytes(data).decode("utf-8"))
            if not isinstance(o, dict):
                raise ValueError("expected a JSON object")
            for f in ("version", "num_vectors", "dimension", "created_at", "updated_at"):
                if not isinstance(o[f], int) or isinstance(o[f], bool):
                    raise ValueError(f"field {f} is not an integer")
            if o["version"] < 0 or o["version"] >= 2 ** 32 or o["num_vectors"] < 0 or o["dimension"] < 0:
                raise ValueError("negative or oversized unsigned field")
            desc = o.get("description")  # Option<String>: a missing key deserialises as None
            if desc is not None and not isinstance(desc, str):
                raise ValueError("description is not a string")
            return cls(o["version"], o["num_vectors"], o["dimension"], o["created_at"], o["updated_at"], desc)
        except (KeyError, ValueError, UnicodeDecodeError) as e:
            raise DeserializationError(str(e)) from None


class FileSystemStorage:
    """StorageBackend for the local filesystem (storage.rs:50-91)."""

    def save(self, path, data: bytes) -> None:
        parent = os.path.dirname(os.fspath(path))
        if parent:
            os.makedirs(parent, exist_ok=True)
        with open(path, "wb") as f:
            f.write(data)

    def load(self, path) -> bytes:
        with open(path, "rb") as f:  # a missing file raises OSError, the reference's Io error
            return f.read()

    def exists(self, path) -> bool:
        return os.path.exists(path)

    def delete(self, path) -> None:
        if os.path.exists(path):  # deleting a missing file is not an error (storage.rs:84-89)
            os.remove(path)


class IndexWriter:
    """storage.rs:93-131.  `IndexWriter.create(path)` or `IndexWriter(stream)`."""

    def __init__(self, stream):
        self._w = stream

    @classmethod
    def create(cls, path) -> "IndexWriter":
        parent = os.path.dirname(os.fspath(path))
        if parent:
            os.makedirs(parent, exist_ok=True)
        return cls(open(path, "wb"))

    def write_metadata(self, metadata: IndexMetadata) -> None:
        self.write_chunk(b"META", metadata.to_json())

    def write_chunk(self, tag: bytes, data: bytes) -> None:
        if len(tag) != 4:
            raise SerializationError("chunk tag must be 4 bytes")
        self._w.write(tag)
        self._w.write(struct.pack("<Q", len(data)))
        self._w.write(data)

    def close(self) -> None:
        self._w.flush()
        if not isinstance(self._w, io.BytesIO):
            self._w.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class IndexReader:
    """storage.rs:133-173.  `IndexReader.open(path)` or `IndexReader(stream)`."""

    def __init__(self, stream):
        self._r = stream

    @classmethod
    def open(cls, path) -> "IndexReader":
        return cls(open(path, "rb"))

    def _exact(self, n: int) -> bytes:
        b = self._r.read(n)
        if len(b) != n:
            raise DeserializationError("failed to fill whole buffer")  # read_exact's UnexpectedEof
        return b

    def read_chunk(self):
        tag = self._exact(4)
        (length,) = struct.unpack("<Q", self._exact(8))
        return tag, self._exact(length)

    def read_metadata(self) -> IndexMetadata:
        tag, data = self.read_chunk()
        if tag != b"META":
            raise DeserializationError("expected META chunk")
        return IndexMetadata.from_json(data)

    def close(self) -> None:
        if not isinstance(self._r, io.BytesIO):
            self._r.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
