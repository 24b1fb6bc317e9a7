"""Pure ops, mirroring sleap_nn/inference/ops/: peaks, crops, paf."""
