"""B200-native fuzzy-match hot path (FuzzyMatch::match of SYSTRAN/fuzzy-match) behind a C ABI.

  fuzzy_match_b200.capi         ctypes binding of include/fuzzy_match_b200.h (libfm_b200.so)
  fuzzy_match_b200.FuzzyMatch   host-side mirror of the reference's fuzzy::FuzzyMatch for the
                                add_tm(Tokens) / sort() / match(Tokens, ...) path
  fuzzy_match_b200.sharded      sentence-id sharded matching over torch.distributed (NCCL)
  fuzzy_match_b200.synth        deterministic synthetic TMs / queries for tests and bench

There is no CPU fallback: every entry point raises if libfm_b200.so is missing or CUDA fails.
"""
from .capi import (FuzzyMatchError, Index, MATCH_DTYPE, WIRE_DTYPE, Params, build_library, library_path,  # noqa: F401
                   load_library)
from .fuzzy_match import ContrastReduce, EditCosts, FuzzyMatch, Match  # noqa: F401
