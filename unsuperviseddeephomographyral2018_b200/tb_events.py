"""Dependency-free TensorBoard event-file writer for the scalars the reference logs every 1000 steps
(code/homography_CNN_synthetic.py:285-293,356-358: `Losses/Learning_rate`, `Losses/Total_{h,rec,ssim,l1,l1_smooth,ncc}_loss`).

An event file is a TFRecord stream — per record: uint64 length, masked CRC32C of the length, payload, masked CRC32C of
the payload — of serialized `Event` protos {wall_time (1, double), step (2, int64), file_version (3, string) |
summary (5) {value (1) {tag (1, string), simple_value (2, float)}}}.  CRC32C comes from tf_checkpoint.py."""
import os
import socket
import struct
import time

from .tf_checkpoint import _field, _varint, crc32c, mask_crc


class SummaryWriter(object):
    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        self.path = os.path.join(log_dir, "events.out.tfevents.%010d.%s" % (int(time.time()), socket.gethostname()))
        self._f = open(self.path, "wb")
        self._record(_field(1, 1, struct.pack("<d", time.time())) + _field(3, 2, _varint(13) + b"brain.Event:2"))

    def _record(self, payload):
        hdr = struct.pack("<Q", len(payload))
        self._f.write(hdr + struct.pack("<I", mask_crc(crc32c(hdr))) + payload + struct.pack("<I", mask_crc(crc32c(payload))))
        self._f.flush()

    def add_scalars(self, scalars, global_step):
        """scalars: {tag: float} -> one Event holding a Summary with one Value per tag."""
        vals = b""
        for tag, v in scalars.items():
            t = tag.encode()
            val = _field(1, 2, _varint(len(t)) + t) + _field(2, 5, struct.pack("<f", float(v)))
            vals += _field(1, 2, _varint(len(val)) + val)
        ev = _field(1, 1, struct.pack("<d", time.time())) + _field(2, 0, _varint(int(global_step))) + _field(5, 2, _varint(len(vals)) + vals)
        self._record(ev)

    def close(self):
        self._f.close()


def read_events(path):
    """-> [(step, {tag: value})] (verifies the record CRCs); used by the tests."""
    from .tf_checkpoint import _parse_message, unmask_crc
    out = []
    with open(path, "rb") as f:
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                break
            (n,) = struct.unpack("<Q", hdr)
            (c,) = struct.unpack("<I", f.read(4))
            assert unmask_crc(c) == crc32c(hdr), "bad length CRC"
            payload = f.read(n)
            (c,) = struct.unpack("<I", f.read(4))
            assert unmask_crc(c) == crc32c(payload), "bad payload CRC"
            ev = _parse_message(payload)
            if 5 in ev:
                scal = {}
                for v in _parse_message(ev[5][0]).get(1, []):
                    m = _parse_message(v)
                    scal[m[1][0].decode()] = struct.unpack("<f", struct.pack("<I", m[2][0]))[0]
                out.append((ev.get(2, [0])[0], scal))
    return out
