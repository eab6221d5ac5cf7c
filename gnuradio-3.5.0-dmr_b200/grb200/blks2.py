"""gnuradio.blks2 as far as the hot path goes: the hier block users actually instantiate.

blks2.pfb_channelizer_ccf (gnuradio-core/src/python/gnuradio/blks2impl/pfb_channelizer.py:25-80) wires
gr.stream_to_streams -> gr.pfb_channelizer_ccf -> gr.vector_to_streams around the channelizer so that ONE interleaved
stream goes in and one stream per channel comes out, and designs the prototype filter itself when none is given.  On
the GPU the two fan-out blocks are addressing: the interleaved stream is what grcuda_pfb_channelizer_ccf_work_interleaved
takes, and the [time][channel] rows it returns are the channel streams side by side.
"""
import numpy as np

from . import blocks, optfir


class pfb_channelizer_ccf:
    """pfb_channelizer_ccf(numchans, taps=None, oversample_rate=1, atten=100): same constructor as the reference's hier
    block (pfb_channelizer.py:32); taps=None designs optfir.low_pass(1, numchans, 0.4, 0.6, ripple, atten) with the
    reference's retry rule (:43-59)."""

    def __init__(self, numchans, taps=None, oversample_rate=1, atten=100):
        self._numchans = int(numchans)
        self._oversample_rate = oversample_rate
        self._taps = np.asarray(taps if taps is not None else optfir.pfb_channelizer_default_taps(self._numchans, atten))
        self.pfb = blocks.pfb_channelizer_ccf(self._numchans, self._taps.astype(np.float32), float(oversample_rate))
        self._started = False

    def taps(self):
        return self._taps

    def set_taps(self, taps):
        self._taps = np.asarray(taps)
        self.pfb.set_taps(self._taps.astype(np.float32))
        self._started = False

    def run(self, x):
        """vector_source(x) -> this block -> numchans vector_sinks: x is the interleaved wideband stream; returns the list
        of numchans channel streams (complex64), exactly what the reference hier block's outputs carry."""
        M = self._numchans
        x = np.ascontiguousarray(x, np.complex64)
        rows = len(x) // M
        T = self.pfb.taps_per_filter()
        om = self.pfb.output_multiple()
        osr = float(self._oversample_rate)
        nout = int(rows * osr) // om * om
        inter = np.concatenate([np.zeros((T, M), np.complex64), x[: rows * M].reshape(rows, M)])   # history rows first
        if not self._started:                 # "history requirements may have changed": the first call returns 0
            y0, _ = self.pfb.general_work_interleaved(nout, inter)
            assert len(y0) == 0
            self._started = True
        y, _ = self.pfb.general_work_interleaved(nout, inter)
        return [np.ascontiguousarray(y[:, c]) for c in range(M)]
