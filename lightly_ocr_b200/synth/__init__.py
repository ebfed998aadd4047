"""Synthetic inputs of the BASELINE.json configs: receipts, crops, score maps and deterministic checkpoints.

Data generators only (no algorithm of the path): used by bench.py, __graft_entry__.smoke() and the tests as the
common input of the CUDA path and of the CPU oracle."""
