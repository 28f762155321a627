"""Synthetic workloads (vocabulary stand-ins + text batches) for tests and bench.py."""
