"""TF-free TFRecord / .size reading: framing, tf.train.Example subset, the feeder contract on a
serialized dataset, and the crop / mirror draws in the reference's RNG order (dataset_.py:444-461,481-501,571-577)."""
import os
import random
import struct
import types

import numpy as np
import pytest

import vlb200  # noqa: F401
from vlb200 import tfrecord
from vlb200.defs import defs
from vlb200.feeder import Dataset


def test_crc32c_known_answers():
    # RFC 3720 test vectors for CRC32C (Castagnoli)
    assert tfrecord.crc32c(b"") == 0x00000000
    assert tfrecord.crc32c(b"123456789") == 0xE3069283
    assert tfrecord.crc32c(bytes(32)) == 0x8A9136AA
    assert tfrecord.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43


def test_example_wire_format_known_bytes():
    """Hand-assembled tf.train.Example bytes (protobuf wire format) parse to the expected features, and the encoder
    reproduces them."""
    # Feature{int64_list{value: [3]}} = 1a 03 0a 01 03 ; map entry {key "a", value}: 0a 01 61 12 05 <feature>
    entry = b"\x0a\x01a\x12\x05\x1a\x03\x0a\x01\x03"
    features = b"\x0a" + bytes([len(entry)]) + entry
    example = b"\x0a" + bytes([len(features)]) + features
    assert tfrecord.parse_example(example) == {"a": [3]}
    assert tfrecord.encode_example({"a": [3]}) == example
    # unpacked repeated int64 (older writers) and negative values
    unpacked = b"\x0a\x01b\x12\x06\x1a\x04\x08\x05\x08\x07"
    feats = b"\x0a" + bytes([len(unpacked)]) + unpacked
    assert tfrecord.parse_example(b"\x0a" + bytes([len(feats)]) + feats) == {"b": [5, 7]}
    assert tfrecord.parse_example(tfrecord.encode_example({"n": [-1, 1 << 40]})) == {"n": [-1, 1 << 40]}


def _write_dataset(tmp_path, cpv, fpc, shape, seed=0, labels=None):
    rng = np.random.default_rng(seed)
    base = str(tmp_path / "data")
    frames, labs = [], []
    with open(base + ".tfrecord", "wb") as f:
        for v, c in enumerate(cpv):
            lab = [int(labels[v])] if labels is not None else [int(rng.integers(0, 7))]
            for _ in range(c * fpc):
                img = rng.integers(0, 256, size=shape, dtype=np.uint8)
                frames.append(img)
                labs.append(lab)
                tfrecord.write_record(f, tfrecord.serialize_frame(img, lab))
    tfrecord.write_size_file(base + ".size", cpv, fpc)
    return base, np.stack(frames), labs


def test_record_roundtrip_and_corruption(tmp_path):
    base, frames, labs = _write_dataset(tmp_path, [1, 2], 2, (5, 6, 3))
    recs = list(tfrecord.read_records(base + ".tfrecord", verify="full"))
    assert len(recs) == 6
    for r, f, l in zip(recs, frames, labs):
        img, lab = tfrecord.deserialize_frame(r)
        assert np.array_equal(img, f) and lab == l
    meta = tfrecord.read_size_file(base + ".size")
    assert meta == dict(items=2, type="video", cpi=[1, 2], fpc=2, labelcount=1)
    # header framing: uint64 length + masked crc
    raw = open(base + ".tfrecord", "rb").read()
    (length,) = struct.unpack("<Q", raw[:8])
    assert length == len(recs[0]) and struct.unpack("<I", raw[8:12])[0] == tfrecord.masked_crc(raw[:8])
    bad = bytearray(raw)
    bad[20] ^= 0xFF
    open(base + ".tfrecord", "wb").write(bytes(bad))
    with pytest.raises(Exception):
        list(tfrecord.read_records(base + ".tfrecord", verify="full"))


def test_size_file_run_length_form(tmp_path):
    p = tmp_path / "x.size"
    p.write_text("items\t5\ntype\tvideo\ncpi\t[(3, 2), (2, 1)]\nfpc\t16\nlabelcount\t1\n")  # serialize.py:144-148
    assert tfrecord.read_size_file(str(p))["cpi"] == [2, 2, 2, 1, 1]


def _opts(base, shape, imgproc, raw=None, fpc=2):
    return types.SimpleNamespace(name="t", data_format=defs.data_format.tfrecord, data_path=base, image_shape=shape,
                                 num_frames_per_clip=fpc, imgproc=imgproc, raw_image_shape=raw, verify_records="full",
                                 mean_image=None)


def test_feeder_contract_on_tfrecord(tmp_path):
    cpv, fpc = [2, 1, 3], 2
    base, frames, labs = _write_dataset(tmp_path, cpv, fpc, (8, 9, 3), labels=[4, 0, 6])
    ds = Dataset(_opts(base, (8, 9, 3), []), batch_size=2, num_classes=7, epochs=1, save_freq_per_epoch=1)
    assert (ds.num_items, ds.num_batches, ds.clips_per_video) == (3, 2, cpv)
    f0, onehot0, cpv0 = ds.next_batch()
    assert cpv0 == [2, 1] and f0.shape == (6, 8, 9, 3) and np.array_equal(f0, frames[:6])
    assert onehot0.shape == (3, 7) and onehot0.argmax(1).tolist() == [4, 4, 0]  # one row per clip
    f1, onehot1, cpv1 = ds.next_batch()
    assert cpv1 == [3] and np.array_equal(f1, frames[6:]) and onehot1.argmax(1).tolist() == [6, 6, 6]
    assert ds.last_crops is not None and not ds.last_crops.any()
    ds.rewind()
    ds.fast_forward(1)  # resume inside the epoch: the first batch's records are skipped
    f1b, _, _ = ds.next_batch()
    assert np.array_equal(f1b, f1)


def test_crop_and_mirror_draws_follow_the_reference_rng_order(tmp_path):
    """dataset_.py:571-577 (admissible offsets range(0, raw - net - 1)), :444-461 (choice(h) then choice(w)),
    :498-500 (`if not randrange(2)` mirrors)."""
    base, frames, _ = _write_dataset(tmp_path, [2], 2, (12, 14, 3))
    ds = Dataset(_opts(base, (8, 9, 3), [defs.imgproc.rand_crop, defs.imgproc.rand_mirror], raw=(12, 14, 3)),
                 batch_size=1, num_classes=7, epochs=1, save_freq_per_epoch=1)
    random.seed(123)
    _, _, _ = ds.next_batch()
    got = ds.last_crops.copy()
    random.seed(123)
    crop_h, crop_w = list(range(0, 12 - 8 - 1)), list(range(0, 14 - 9 - 1))
    exp = []
    for _ in range(4):
        hh = random.choice(crop_h)
        ww = random.choice(crop_w)
        exp.append((hh, ww, 1 if not random.randrange(2) else 0))
    assert got.tolist() == [list(e) for e in exp]
    dc = Dataset(_opts(base, (8, 9, 3), [defs.imgproc.center_crop], raw=(12, 14, 3)), batch_size=1, num_classes=7,
                 epochs=1, save_freq_per_epoch=1)
    dc.next_batch()
    assert dc.last_crops.tolist() == [[2, 2, 0]] * 4  # floor((12-8)/2), floor((14-9)/2)


def test_read_time_resize_modes_keep_frames_at_their_serialized_size(tmp_path):
    """imgproc raw_resize / resize (dataset_.py:481-491): the host hands over the serialized frames untouched together
    with the target of the device-side imresize; crops are drawn against raw_image_shape; resize loses against a crop
    (`elif`), and a double resampling is refused."""
    base, frames, _ = _write_dataset(tmp_path, [1, 2], 2, (6, 7, 3))
    rr = Dataset(_opts(base, (8, 9, 3), [defs.imgproc.raw_resize, defs.imgproc.center_crop], raw=(12, 14, 3)),
                 batch_size=2, num_classes=7, epochs=1, save_freq_per_epoch=1)
    assert rr.resize_to == (12, 14)
    f, onehot, cpv = rr.next_batch()
    assert f.shape == (6, 6, 7, 3) and np.array_equal(f, frames) and rr.stored_shape == (6, 7, 3)
    assert rr.last_crops.tolist() == [[2, 2, 0]] * 6  # centre crop of the 12 x 14 resampled frame to 8 x 9
    rs = Dataset(_opts(base, (8, 9, 3), [defs.imgproc.resize]), batch_size=2, num_classes=7, epochs=1,
                 save_freq_per_epoch=1)
    assert rs.resize_to == (8, 9)
    f2, _, _ = rs.next_batch()
    assert f2.shape == (6, 6, 7, 3) and not rs.last_crops.any()
    # crop wins over resize: frames must then already have raw_image_shape
    rc = Dataset(_opts(base, (4, 5, 3), [defs.imgproc.resize, defs.imgproc.center_crop], raw=(6, 7, 3)), batch_size=2,
                 num_classes=7, epochs=1, save_freq_per_epoch=1)
    assert rc.resize_to is None
    rc.next_batch()
    assert rc.last_crops.tolist() == [[1, 1, 0]] * 6
    with pytest.raises(Exception, match="twice"):
        Dataset(_opts(base, (8, 9, 3), [defs.imgproc.raw_resize, defs.imgproc.resize], raw=(12, 14, 3)), batch_size=2,
                num_classes=7, epochs=1, save_freq_per_epoch=1)
    with pytest.raises(Exception, match="raw_image_shape"):
        Dataset(_opts(base, (8, 9, 3), [defs.imgproc.raw_resize]), batch_size=2, num_classes=7, epochs=1,
                save_freq_per_epoch=1)
