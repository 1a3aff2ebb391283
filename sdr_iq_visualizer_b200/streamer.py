"""SDR data streamer -- drop-in for /root/reference/app/sdr/streamer.py with the spectrum computed
on the GPU.

Kept from the reference: the ``SDRDataStreamer`` constructor and its
defaults (streamer.py:8-10), attributes ``uri/sample_rate/center_freq/rx_lo/rx_rf_bandwidth/
rx_buffer_size/sdr/data_queue/running/thread/connected``, the methods ``connect / is_connected /
start_streaming / stop_streaming / reconnect / get_status / get_latest_data``, the module global
``sdr_streamer`` and the result dict of one frame (streamer.py:123-130):
``{'time','samples','freqs','power_db','sample_rate','center_freq'}`` with ``freqs``/``power_db``
float64[N] in fftshift order.  ``import adi`` stays at module top so tests can mock it.

Changed on purpose:
  * the three numpy lines (streamer.py:119-121) are one call into ``spectral.stream_frame`` (CUDA);
    the frequency axis is cached per (N, fs, fc) instead of being rebuilt per buffer;
  * any exception raised after ``rx()`` returned (``SpectralError``, but also a ``ValueError`` or ``TypeError``
    out of the processing) is a COMPUTE fault: logged rate-limited, counted separately, backed off and, after
    ``COMPUTE_FAULT_LIMIT`` in a row, the stream stops with ``compute_state == 'failed'`` -- it never looks like a
    radio fault and never triggers a reconnect (reference :157-174 would);
  * ``get_latest_data()`` keeps the reference's pop-one (oldest first) semantics so that unmodified consumers
    share the queue as before; ``get_newest_data()`` / ``get_latest_data(drain=True)`` discard stale frames and
    ``peek_latest()`` does not consume at all;
  * ``get_status()`` also reports the compute state, compute errors, samples processed and H2D bytes.
"""
import logging
import queue
import threading
import time

import adi
import numpy as np

from . import spectral
from ._native import SpectralError

logger = logging.getLogger(__name__)

_FATAL_ERRNOS = (9, 10054)        # bad descriptor, connection reset  (streamer.py:137)
_RECONNECT_ERRNOS = (110, 113)    # timed out, host unreachable       (streamer.py:148)


class SDRDataStreamer:
    def __init__(self, uri="ip:192.168.2.1", sample_rate=1_000_000,
                 center_freq=2_400_000_000, rx_lo=2_400_000_000,
                 rx_rf_bandwidth=4_000_000, rx_buffer_size=2**12):
        self.uri = uri
        self.sample_rate = sample_rate
        self.center_freq = center_freq
        self.rx_lo = rx_lo
        self.rx_rf_bandwidth = rx_rf_bandwidth
        self.rx_buffer_size = rx_buffer_size
        self.sdr = None
        self.data_queue = queue.Queue(maxsize=100)
        self.running = False
        self.thread = None
        self.connected = False
        self._reconnect_lock = threading.Lock()
        self._axis_key = None
        self._axis = None
        self.compute_errors = 0
        self.samples_processed = 0
        self.h2d_bytes = 0
        self._latest = None
        self.waterfall_range = None   # (vmin_db, vmax_db): frames then also carry the uint8 waterfall row ('wf_row')

    # ------------------------------------------------------------------ radio control (host I/O)
    def connect(self):
        """Open the Pluto and push the RX configuration; True on success."""
        try:
            self.sdr = adi.Pluto(uri=self.uri)
            self.sdr.sample_rate = int(self.sample_rate)
            self.sdr.rx_rf_bandwidth = int(self.rx_rf_bandwidth)
            self.sdr.rx_lo = int(self.rx_lo)
            self.sdr.rx_buffer_size = self.rx_buffer_size
            logger.info("Connected to SDR at %s (fs=%s, lo=%s, bw=%s)", self.uri, self.sdr.sample_rate,
                        self.sdr.rx_lo, self.sdr.rx_rf_bandwidth)
            self.connected = True
        except Exception as exc:
            logger.error("Failed to connect to SDR: %s", exc)
            self.sdr = None
            self.connected = False
        return self.connected

    def is_connected(self):
        """Cheap check; never touches device properties (that would disturb rx())."""
        return self.sdr is not None and self.connected

    def start_streaming(self):
        if not self.sdr:
            logger.error("SDR not connected")
            return False
        self.running = True
        self.thread = threading.Thread(target=self._stream_data, daemon=True)
        self.thread.start()
        return True

    def stop_streaming(self):
        self.running = False
        if self.thread:
            self.thread.join(timeout=1)
        logger.info("Stopped SDR data streaming")

    def reconnect(self):
        logger.info("Attempting to reconnect to SDR...")
        self.sdr = None
        return self.connect()

    def _attempt_reconnect(self, max_attempts=5, base_delay=0.5):
        """Exponential-backoff reconnect under a lock; True on success."""
        with self._reconnect_lock:
            for attempt in range(1, max_attempts + 1):
                if self.reconnect():
                    logger.info("Auto-reconnect succeeded on attempt %d.", attempt)
                    return True
                time.sleep(min(base_delay * (2 ** (attempt - 1)), 5.0))
        return False

    # ------------------------------------------------------------------ block mode: pinned-ring ingest
    def enable_block_mode(self, nfft=4096, overlap=0.75, window="hann", frames_per_block=1024, n_slots=4,
                          vmin=0.0, vmax=130.0, device=0):
        """High-rate path next to the per-buffer dicts: every rx buffer is also written, as raw int16 I,Q, into a
        page-locked ring slot; a full slot (``frames_per_block`` hops) goes H2D -> fused STFT -> D2H on three
        streams while the next slot fills.  Each block yields its uint8 waterfall rows, Welch PSD, max-hold and the
        classifier measurements (computed on the device behind the STFT kernel); ``get_latest_block()`` returns the
        newest one, ``get_status()`` reports the PCIe byte counters.  The reference queues one dict per buffer
        instead (streamer.py:123-131,186-194)."""
        from .ring import StreamRing
        hop = max(1, int(round(nfft * (1.0 - overlap))))
        self._blk_plan = spectral.SpectralPlan(nfft, hop, window, spectral.FMT_CI16, device=device)
        self._blk_ring = StreamRing(self._blk_plan, n_slots=n_slots, slot_samples=frames_per_block * hop, wf_rows=True,
                                    welch=True, maxhold=True, vmin=vmin, vmax=vmax, features=True,
                                    sample_rate=float(self.sample_rate))
        self._blk_cfg = {"nfft": nfft, "hop": hop, "vmin": vmin, "vmax": vmax, "slot_samples": frames_per_block * hop}
        self._blk_slot, self._blk_fill, self._blk_pending = None, 0, 0
        self._blk_maxhold = np.zeros(nfft, np.float32)
        self._blk_axis = spectral.freq_axis(nfft, self.sample_rate, self.center_freq)
        self._latest_block = None
        self.blocks_done = 0
        return self

    def disable_block_mode(self):
        ring, plan = getattr(self, "_blk_ring", None), getattr(self, "_blk_plan", None)
        self._blk_ring = self._blk_plan = None
        if ring is not None:
            ring.close()
        if plan is not None:
            plan.close()

    def _publish_block(self):
        """Collect the oldest finished slot (blocks only until ITS copies are done) and publish it."""
        from . import classifier
        blk = self._blk_ring.collect()
        np.maximum(self._blk_maxhold, blk["maxhold"], out=self._blk_maxhold)        # running max-hold across blocks
        out = {"time": time.time(), "seq": blk["seq"], "first_frame": blk["first_frame"], "n_frames": blk["n_frames"],
               "freqs": self._blk_axis, "wf_rows": blk["wf_rows"].copy(), "welch_acc": blk["welch_acc"].copy(),
               "maxhold": self._blk_maxhold.copy(), "pxx_db": None if blk["pxx_db"] is None else blk["pxx_db"].copy(),
               "features": None, "classification": None, "vmin": self._blk_cfg["vmin"], "vmax": self._blk_cfg["vmax"]}
        if blk["features"] is not None:
            m = classifier.with_hz(self._blk_axis, blk["features"])
            out["features"] = m
            out["classification"] = classifier.classify_from_measurements(self._blk_axis, m, self._blk_cfg["nfft"])
        self._blk_ring.release()
        self._blk_pending -= 1
        self._latest_block = out
        self.blocks_done += 1

    def _feed_ring(self, samples):
        """Append one rx buffer to the ring as interleaved int16 (pyadi-iio hands back the raw integer counts as
        complex128, streamer.py:114, so the conversion is exact)."""
        x = np.asarray(samples)
        iq = np.empty(2 * x.size, np.int16)
        iq[0::2] = np.clip(x.real, -32768, 32767)
        iq[1::2] = np.clip(x.imag, -32768, 32767)
        cap = self._blk_cfg["slot_samples"]
        pos = 0
        while pos < x.size:
            if self._blk_slot is None:
                if self._blk_pending >= self._blk_ring_slots() - 1:
                    self._publish_block()
                self._blk_slot, self._blk_fill = self._blk_ring.acquire(), 0
            take = min(x.size - pos, cap - self._blk_fill)
            self._blk_slot[2 * self._blk_fill: 2 * (self._blk_fill + take)] = iq[2 * pos: 2 * (pos + take)]
            self._blk_fill += take
            pos += take
            if self._blk_fill == cap:
                self._blk_ring.commit(cap)
                self._blk_slot = None
                self._blk_pending += 1
                if self._blk_pending >= 2:       # keep one block in flight, publish the one before it
                    self._publish_block()
        self.h2d_bytes += 4 * x.size

    def _blk_ring_slots(self):
        return int(self._blk_ring._n_slots)

    def flush_blocks(self):
        """Commit a partly filled slot and publish everything in flight (end of a capture)."""
        if getattr(self, "_blk_ring", None) is None:
            return
        if self._blk_slot is not None and self._blk_fill:
            self._blk_ring.commit(self._blk_fill)
            self._blk_pending += 1
        self._blk_slot = None
        while self._blk_pending:
            self._publish_block()

    def get_latest_block(self):
        """Newest finished block of the ring ingest, or None (``enable_block_mode`` first)."""
        return getattr(self, "_latest_block", None)

    # ------------------------------------------------------------------ the hot path
    def _frequency_axis(self, n):
        key = (n, self.sample_rate, self.center_freq)
        if key != self._axis_key:
            self._axis = spectral.freq_axis(n, self.sample_rate, self.center_freq)
            self._axis_key = key
        return self._axis

    def process_buffer(self, samples):
        """One rx buffer -> the frame dict of streamer.py:123-130 (spectrum on the GPU)."""
        n = len(samples)
        wf_row = None
        if self.waterfall_range is None:
            _, power_db = spectral.stream_frame(samples, self.sample_rate, self.center_freq)
        else:
            _, power_db, wf_row = spectral.stream_frame(samples, self.sample_rate, self.center_freq,
                                                        wf_range=self.waterfall_range)
        self.samples_processed += n
        self.h2d_bytes += n * 8
        if getattr(self, "_blk_ring", None) is not None:
            self._feed_ring(samples)
        extra = {} if wf_row is None else {'wf_row': wf_row}
        return {
            **extra,
            'time': time.time(),
            'samples': samples,
            'freqs': self._frequency_axis(n),
            'power_db': power_db,
            'sample_rate': self.sample_rate,
            'center_freq': self.center_freq,
        }

    def _stream_data(self):
        """Read buffers until stopped; radio errors back off / reconnect, compute errors do not."""
        radio_errors = 0
        backoff, max_backoff = 0.1, 1.6
        self.last_success_ts = None
        self.total_frames = 0

        def recovered(attempts, delay):
            nonlocal radio_errors, backoff
            if self._attempt_reconnect(max_attempts=attempts, base_delay=delay):
                radio_errors, backoff = 0, 0.1
                return True
            return False

        while self.running:
            if not self.is_connected():
                logger.warning("SDR not connected; trying auto-reconnect before stopping...")
                if not recovered(5, 0.5):
                    logger.error("Auto-reconnect failed; stopping stream.")
                    self.running = False
                    break
            try:
                samples = self.sdr.rx()  # blocking hardware read
                radio_errors, backoff = 0, 0.1
                # everything after rx() is compute: ANY exception from it (SpectralError, but also a ValueError /
                # TypeError / KeyError out of process_buffer or the ring feed) is a compute fault, never a radio fault
                try:
                    frame = self.process_buffer(samples)
                except Exception as exc:
                    if not self._compute_fault(exc):
                        break
                    continue
                self._compute_ok()
                self._push(frame)
                self.last_success_ts = frame['time']
                self.total_frames += 1
            except OSError as exc:
                radio_errors += 1
                err = getattr(exc, 'errno', None)
                if err in _FATAL_ERRNOS:
                    logger.error("Fatal OS error errno=%s; attempting auto-reconnect.", err)
                    self.connected = False
                    if recovered(5, 0.5):
                        continue
                    logger.error("Auto-reconnect failed after fatal error; stopping stream.")
                    self.running = False
                    break
                logger.error("Non-fatal OS error reading SDR (errno=%s): %s", err, exc)
                if err in _RECONNECT_ERRNOS and recovered(3, 0.2):
                    continue
            except Exception as exc:
                radio_errors += 1
                logger.error("Error reading SDR data: %s", exc)

            if radio_errors:
                backoff = min(backoff * 2, max_backoff)
                logger.warning("Read error #%d; backoff %.2fs", radio_errors, backoff)
                time.sleep(backoff)
                if radio_errors >= 3:
                    logger.warning("Too many consecutive errors; attempting auto-reconnect.")
                    self.connected = False
                    if recovered(5, 0.5):
                        continue
                    logger.error("Auto-reconnect failed after repeated errors; stopping stream.")
                    self.running = False
                    break

    # ------------------------------------------------------------------ compute-fault handling
    COMPUTE_FAULT_LIMIT = 25        # consecutive failures after which the stream stops ("failed")
    COMPUTE_BACKOFF_MAX_S = 1.6

    def _compute_fault(self, exc) -> bool:
        """Count a compute failure, log it rate-limited, back off; returns False when the stream must stop."""
        self.compute_errors += 1
        self.compute_errors_consecutive = getattr(self, 'compute_errors_consecutive', 0) + 1
        k = self.compute_errors_consecutive
        self.compute_last_error = f"{type(exc).__name__}: {exc}"
        self.compute_state = 'degraded'
        if k == 1 or k % 10 == 0:
            logger.error("GPU spectrum failed (not a radio fault; %d in a row): %s", k, self.compute_last_error)
        if k >= self.COMPUTE_FAULT_LIMIT:
            logger.error("GPU spectrum failed %d times in a row; stopping the stream (radio left connected).", k)
            self.compute_state = 'failed'
            self.running = False
            return False
        time.sleep(min(0.05 * (2 ** min(k - 1, 5)), self.COMPUTE_BACKOFF_MAX_S))
        return True

    def _compute_ok(self) -> None:
        if getattr(self, 'compute_errors_consecutive', 0):
            logger.info("GPU spectrum recovered after %d failures.", self.compute_errors_consecutive)
        self.compute_errors_consecutive = 0
        self.compute_state = 'ok'

    # ------------------------------------------------------------------ queue + status
    def get_status(self):
        last = getattr(self, 'last_success_ts', None)
        return {
            'connected': self.connected,
            'running': self.running,
            'queue_size': self.data_queue.qsize(),
            'last_success_age_ms': (time.time() - last) * 1000 if last else None,
            'total_frames': getattr(self, 'total_frames', 0),
            'compute_errors': self.compute_errors,
            'compute_state': getattr(self, 'compute_state', 'ok'),          # ok | degraded (backing off) | failed (stopped)
            'compute_errors_consecutive': getattr(self, 'compute_errors_consecutive', 0),
            'compute_last_error': getattr(self, 'compute_last_error', None),
            'samples_processed': self.samples_processed,
            'h2d_bytes': self.h2d_bytes,
            'blocks_done': getattr(self, 'blocks_done', 0),
            'ring': self._blk_ring.stats() if getattr(self, '_blk_ring', None) is not None else None,
        }

    def _push(self, data):
        """Bounded queue, drop-oldest when full (streamer.py:186-194)."""
        self._latest = data
        while True:
            try:
                self.data_queue.put_nowait(data)
                return
            except queue.Full:
                try:
                    self.data_queue.get_nowait()
                except queue.Empty:
                    return

    def get_latest_data(self, drain: bool = False):
        """Reference semantics by default (streamer.py:196-200): pop ONE frame, the oldest queued, or None when the
        queue is empty -- unmodified consumers (dashboard tick callbacks.py:104, recorder callbacks.py:263, chatbot
        chatbot.py:149) share this call and must not starve each other.  ``drain=True`` discards everything but the
        newest frame and returns it (see also ``get_newest_data`` / ``peek_latest``)."""
        if not drain:
            try:
                return self.data_queue.get_nowait()
            except queue.Empty:
                return None
        latest = None
        while True:
            try:
                latest = self.data_queue.get_nowait()
            except queue.Empty:
                return latest

    def get_newest_data(self):
        """Newest queued frame (older ones are discarded), or None."""
        return self.get_latest_data(drain=True)

    def peek_latest(self):
        """Newest frame WITHOUT consuming the queue: a second consumer (the chatbot's classify tool,
        /root/reference/app/chatbot/chatbot.py:149) no longer races the dashboard tick for queue items."""
        return self._latest


# Shared global instance (streamer.py:203)
sdr_streamer = SDRDataStreamer()
