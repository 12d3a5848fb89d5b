"""salient_plusplus_b200 -- B200 (sm_100a) native mini-batch generation for SALIENT++-style
distributed GNN training: neighbour sampling, dedup/relabel, partition-book translation,
VIP-cache split, feature gather and the NVLink peer-to-peer miss fetch, behind the reference's
``fast_sampler`` / ``fast_trainer`` Python API.

Importing the package does not load the CUDA library; the first call into
``salient_plusplus_b200.fast_sampler`` does, and raises if it has not been built.
"""
__version__ = "0.1.0"
