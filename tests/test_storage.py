"""META chunk container (src/core/storage.rs:93-173) and IndexMetadata (storage.rs:15-48): the cases of
the reference's own tests (storage.rs:186-375) plus known-answer bytes for the chunk framing and the
serde_json field order.  Host-side format code: no GPU."""
import io
import json
import struct

import pytest

from islands_b200.storage import (DeserializationError, FileSystemStorage, IndexMetadata, IndexReader, IndexWriter)


def test_filesystem_storage(tmp_path):  # storage.rs:186-201
    path = tmp_path / "test.bin"
    s = FileSystemStorage()
    s.save(path, b"test data")
    assert s.exists(path)
    assert s.load(path) == b"test data"
    s.delete(path)
    assert not s.exists(path)


def test_filesystem_storage_nested_and_missing(tmp_path):  # storage.rs:203-233
    s = FileSystemStorage()
    path = tmp_path / "nested/deep/path/test.bin"
    s.save(path, b"nested data")
    assert s.load(path) == b"nested data"
    s.delete(tmp_path / "nonexistent.bin")  # no error
    with pytest.raises(OSError):
        s.load(tmp_path / "nonexistent.bin")


def test_index_metadata_new_and_json():  # storage.rs:235-259
    meta = IndexMetadata.new(100, 128)
    assert meta.version == IndexMetadata.CURRENT_VERSION == 1
    assert (meta.num_vectors, meta.dimension) == (100, 128)
    assert meta.created_at > 0 and meta.created_at == meta.updated_at
    assert meta.description is None
    meta = IndexMetadata.new(50, 64)
    meta.description = "test index"
    parsed = IndexMetadata.from_json(meta.to_json())
    assert parsed == meta


def test_metadata_json_is_serde_jsons_bytes():
    """serde_json::to_vec of the struct: declaration order, compact, None -> null, UTF-8 kept raw,
    quotes / backslashes / control characters escaped."""
    m = IndexMetadata(1, 42, 128, 1700000000, 1700000001, None)
    assert m.to_json() == b'{"version":1,"num_vectors":42,"dimension":128,"created_at":1700000000,"updated_at":1700000001,"description":null}'
    m.description = 'isländs "q"\\\n'
    assert m.to_json().endswith('"description":"isländs \\"q\\"\\\\\\n"}'.encode("utf-8"))
    assert IndexMetadata.from_json(m.to_json()) == m
    # a reader accepts any key order and a missing Option field
    assert IndexMetadata.from_json(b'{"dimension":3,"version":1,"updated_at":5,"created_at":4,"num_vectors":2}') == \
        IndexMetadata(1, 2, 3, 4, 5, None)


def test_writer_reader_roundtrip(tmp_path):  # storage.rs:261-322, :358-375
    path = tmp_path / "nested/dir/index.leann"
    original = IndexMetadata.new(42, 128)
    original.description = "test description"
    with IndexWriter.create(path) as w:
        w.write_metadata(original)
        w.write_chunk(b"LEAN", b"\x01\x02\x03")
    raw = path.read_bytes()
    body = original.to_json()
    assert raw == b"META" + struct.pack("<Q", len(body)) + body + b"LEAN" + struct.pack("<Q", 3) + b"\x01\x02\x03"
    with IndexReader.open(path) as r:
        assert r.read_metadata() == original
        assert r.read_chunk() == (b"LEAN", b"\x01\x02\x03")
        with pytest.raises(DeserializationError):
            r.read_chunk()  # end of file


def test_writer_with_cursor():  # storage.rs:324-334
    buf = io.BytesIO()
    IndexWriter(buf).write_metadata(IndexMetadata.new(5, 16))
    buf.seek(0)
    assert IndexReader(buf).read_metadata().num_vectors == 5


def test_reader_rejects_wrong_tag_truncation_and_bad_json():  # storage.rs:336-356
    bad = io.BytesIO(b"BAAD" + struct.pack("<Q", 8) + b"testdata")
    with pytest.raises(DeserializationError, match="expected META chunk"):
        IndexReader(bad).read_metadata()
    with pytest.raises(DeserializationError):
        IndexReader(io.BytesIO(b"META" + struct.pack("<Q", 100) + b"{}")).read_metadata()
    with pytest.raises(DeserializationError):
        IndexReader(io.BytesIO(b"META" + struct.pack("<Q", 2) + b"{}")).read_metadata()
    body = json.dumps({"version": -1, "num_vectors": 1, "dimension": 1, "created_at": 0, "updated_at": 0}).encode()
    with pytest.raises(DeserializationError):
        IndexReader(io.BytesIO(b"META" + struct.pack("<Q", len(body)) + body)).read_metadata()
