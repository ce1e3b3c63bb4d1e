"""B200-native drop-in for MagellanMapper's volumetric blob-detection path.

Sub-packages mirror the reference layout for the accelerated path only:
``cv.detector``, ``cv.stack_detect``, ``cv.chunking``, ``plot.plot_3d``,
``settings.config`` / ``settings.roi_prof``.  All arithmetic runs in the
C-ABI CUDA library built from ``csrc/`` (``include/mmb200.h``); there is no
CPU fallback - calling a compute entry point without the library raises.
"""
__version__ = "0.1.0"
