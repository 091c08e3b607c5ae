"""Mirror of ``tetraear.signal`` for the IQ->bursts path (reference: tetraear/signal/__init__.py:11-28).

Only the classes on the path are provided (plus the scanner's per-dwell detector);
device capture and the scan loop stay with the reference.
"""
from .processor import SignalProcessor
from .spectrum import SpectrumAnalyzer, compute_spectrum, find_peak_near, find_spike_band_near, afc_step
from .scanner import TetraSignalDetector

__all__ = ["SignalProcessor", "SpectrumAnalyzer", "compute_spectrum", "find_peak_near", "find_spike_band_near", "afc_step", "TetraSignalDetector"]
