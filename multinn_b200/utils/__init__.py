"""Host-side utilities around the hot path (mirror reference multinn/utils/): dataset loading, piece batching, streaming
evaluation, training statistics and the epoch loop of train.py."""
