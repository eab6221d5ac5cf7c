"""gr_framer_sink_1, oracle side only: the plain-C restatement of the framer's state machine
against the compiled reference class (with a stand-in message queue) and the committed fixture.  Byte / integer work:
the packets must be identical.  The GPU block for this row is not written yet."""
import numpy as np
import pytest


def split_packets(fx):
    out, pos = [], 0
    for off, n in zip(fx["framer_offsets"], fx["framer_lengths"]):
        out.append((int(off), bytes(fx["framer_payloads"][pos:pos + n])))
        pos += n
    return out


def test_oracle_fixture(orc, golden_next):
    want = split_packets(golden_next)
    assert [len(p[1]) for p in want] == [5, 0, 1, 300, 2, 64]          # packets 3 and 6 had a corrupted header
    for nchunks in (1, 23, 997):
        f, got = orc.Framer(), []
        for chunk in np.array_split(golden_next["framer_stream"], nchunks):
            got += f.work(chunk)
        assert got == want, nchunks


def test_oracle_live_vs_reference(orc, ref):
    rng = np.random.default_rng(17)
    for trial in range(6):
        pk = [(int(rng.integers(0, 16)), bytes(rng.integers(0, 256, int(rng.integers(0, 200))).astype(np.uint8)))
              for _ in range(int(rng.integers(1, 12)))]
        stream = orc.framer_make_stream(rng, pk, gap=(0, 60), corrupt_header_every=int(rng.integers(0, 4)))
        # stray flags inside payloads and gaps must be ignored while a packet is being read and honoured otherwise
        stray = rng.integers(0, len(stream), 5)
        stream[stray] |= 2
        f, o, want, got = ref.FramerSink(), orc.Framer(), [], []
        for chunk in np.array_split(stream, int(rng.integers(1, 40))):
            want += f.work(chunk)
            got += o.work(chunk)
        assert got == want, trial


def test_oracle_known_packet(orc):
    """Hand-made case in the style of the reference's packet_utils framing: offset 9, payload b'GR', flag on the first
    header bit; a header whose halves differ is dropped."""
    def stream(h_hi, h_lo, payload):
        hdr = [(((h_hi << 16) | h_lo) >> (31 - b)) & 1 for b in range(32)]
        bits = np.array([0, 1, 1, 0] + hdr + [(c >> (7 - b)) & 1 for c in payload for b in range(8)] + [0] * 9, np.uint8)
        bits[4] |= 2
        return bits
    h = (9 << 12) | 2
    assert orc.Framer().work(stream(h, h, b"GR")) == [(9, b"GR")]
    assert orc.Framer().work(stream(h, h ^ 1, b"GR")) == []
