"""TensorFlow-free reader (and writer) of the reference's serialized datasets.

The reference stores frames with `tf.python_io.TFRecordWriter` as `tf.train.Example` protobufs with the features
`height`, `width`, `depth`, `label` (int64 lists) and `image_raw` (bytes: the uint8 HWC frame, BGR)
(serialize.py:246-256) and reads them back with `tf.python_io.tf_record_iterator` (dataset_.py:100-133,171-217).
Next to the `.tfrecord` file sits a `.size` text file (serialize.py:138-151, dataset_.py:701-756).

TFRecord framing (public format): uint64 length, uint32 masked-crc32c(length), payload, uint32 masked-crc32c(payload),
all little endian; mask(c) = ((c >> 15 | c << 17) + 0xa282ead8) mod 2^32.
Protobuf wire subset needed: varint (type 0) and length-delimited (type 2) fields.
  Example{1: Features}  Features{1: map<string, Feature>}  map entry{1: key, 2: Feature}
  Feature{1: BytesList, 2: FloatList, 3: Int64List}  BytesList{1: repeated bytes}  Int64List{1: packed varints}
"""
import ast
import struct

import numpy as np

from .utils import error

# ----------------------------------------------------------------------------------------------------------
# crc32c (Castagnoli), table driven
# ----------------------------------------------------------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        table = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            table.append(c)
        _CRC_TABLE = table
    return _CRC_TABLE


def crc32c(data):
    table = _crc_table()
    c = 0xFFFFFFFF
    for b in data:
        c = table[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------------------------------
# record framing
# ----------------------------------------------------------------------------------------------------------
def read_records(path, verify="length"):
    """Yield the payload of every record.  verify: None, "length" (header CRC only, cheap) or "full"."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                error("Truncated TFRecord header in %s" % path)
            (length,) = struct.unpack("<Q", head[:8])
            if verify and struct.unpack("<I", head[8:])[0] != masked_crc(head[:8]):
                error("Corrupt TFRecord length field in %s" % path)
            payload = f.read(length)
            tail = f.read(4)
            if len(payload) < length or len(tail) < 4:
                error("Truncated TFRecord payload in %s" % path)
            if verify == "full" and struct.unpack("<I", tail)[0] != masked_crc(payload):
                error("Corrupt TFRecord payload in %s" % path)
            yield payload


def write_record(f, payload):
    head = struct.pack("<Q", len(payload))
    f.write(head)
    f.write(struct.pack("<I", masked_crc(head)))
    f.write(payload)
    f.write(struct.pack("<I", masked_crc(payload)))


# ----------------------------------------------------------------------------------------------------------
# protobuf subset
# ----------------------------------------------------------------------------------------------------------
def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) of one message; value is an int (varint) or a memoryview slice."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        else:
            error("Unsupported protobuf wire type %d" % wt)
        yield num, wt, val


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_example(payload):
    """tf.train.Example -> {feature name: list of bytes | list of int | list of float}."""
    out = {}
    buf = memoryview(payload)
    for num, _, features in _fields(buf):
        if num != 1:
            continue
        for fnum, _, entry in _fields(features):
            if fnum != 1:
                continue
            key, feat = None, None
            for enum_, _, val in _fields(entry):
                if enum_ == 1:
                    key = bytes(val).decode("utf-8")
                elif enum_ == 2:
                    feat = val
            values = []
            if feat is not None:
                for kind, _, lst in _fields(feat):
                    for vnum, wt, val in _fields(lst):
                        if vnum != 1:
                            continue
                        if kind == 1:
                            values.append(bytes(val))
                        elif kind == 3:
                            if wt == 2:  # packed
                                p = 0
                                while p < len(val):
                                    v, p = _varint(val, p)
                                    values.append(_signed64(v))
                            else:
                                values.append(_signed64(val))
                        elif kind == 2:
                            if wt == 2:
                                values.extend(np.frombuffer(bytes(val), dtype="<f4").tolist())
                            else:
                                values.append(struct.unpack("<f", bytes(val))[0])
            out[key] = values
    return out


def _enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_ld(num, payload):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_example(features):
    """{name: bytes | list of int} -> serialized tf.train.Example (what serialize.py:246-256 writes)."""
    entries = b""
    for name in sorted(features):
        val = features[name]
        if isinstance(val, (bytes, bytearray)):
            feat = _enc_ld(1, _enc_ld(1, bytes(val)))
        else:
            packed = b"".join(_enc_varint(int(v)) for v in val)
            feat = _enc_ld(3, _enc_ld(1, packed))
        entries += _enc_ld(1, _enc_ld(1, name.encode("utf-8")) + _enc_ld(2, feat))
    return _enc_ld(1, entries)


# ----------------------------------------------------------------------------------------------------------
# the reference's frame records and .size files
# ----------------------------------------------------------------------------------------------------------
def deserialize_frame(payload):
    """dataset_.py:100-133: (uint8 image [height, width, depth], label list)."""
    ex = parse_example(payload)
    for k in ("image_raw", "height", "width", "depth", "label"):
        if k not in ex:
            error("TFRecord example without feature `%s`" % k)
    h, w, d = int(ex["height"][0]), int(ex["width"][0]), int(ex["depth"][0])
    img = np.frombuffer(ex["image_raw"][0], dtype=np.uint8)
    if img.size != h * w * d:
        error("image_raw holds %d bytes, expected %dx%dx%d" % (img.size, h, w, d))
    return img.reshape(h, w, d), [int(v) for v in ex["label"]]


def serialize_frame(image, label):
    image = np.ascontiguousarray(image, dtype=np.uint8)
    label = list(label) if isinstance(label, (list, tuple, np.ndarray)) else [int(label)]
    return encode_example({"height": [image.shape[0]], "width": [image.shape[1]], "depth": [image.shape[2]],
                           "label": label, "image_raw": image.tobytes()})


def read_size_file(path):
    """dataset_.py:701-756 / utils_.py:234-243: returns dict(items, type, cpi (expanded list or None), fpc, labelcount)."""
    info_ = {}
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            key, value = line.strip().split("\t")
            info_[key.strip()] = value.strip()
    for k in ("items", "type", "cpi", "fpc", "labelcount"):
        if k not in info_:
            error("Size file %s lacks the entry `%s`" % (path, k))
    items = int(ast.literal_eval(info_["items"]))
    cpi = ast.literal_eval(info_["cpi"])
    fpc = ast.literal_eval(info_["fpc"])
    if isinstance(cpi, list):
        if cpi and isinstance(cpi[0], tuple):  # run-length encoded (count, value) pairs (serialize.py:144-145)
            cpi = [item for num, item in cpi for _ in range(num)]
        if len(cpi) != items:
            error("Read %d items but got cpv list of size %d" % (items, len(cpi)))
    return dict(items=items, type=info_["type"], cpi=cpi, fpc=fpc, labelcount=int(info_["labelcount"]))


def write_size_file(path, clips_per_item, fpc, mode="video", max_num_labels=1):
    import itertools
    with open(path, "w") as f:
        f.write("items\t%d\n" % len(clips_per_item))
        f.write("type\t%s\n" % mode)
        f.write("cpi\t%s\n" % [(len(list(g)), k) for k, g in itertools.groupby(clips_per_item)])
        f.write("fpc\t%s\n" % str(fpc))
        f.write("labelcount\t%s\n" % str(max_num_labels))
