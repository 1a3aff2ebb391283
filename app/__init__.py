"""Drop-in package layout of the reference (`app.processing.*`, `app.sdr.*`): thin re-exports of
sdr_iq_visualizer_b200 so that `from app.processing.classifier import ...` (reference
tests/test_classifier.py:3, app/dashboard/callbacks.py:14) resolves to the B200 implementation."""
