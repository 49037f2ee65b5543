"""exomedepth_b200 — B200-native CNV-calling core behind ExomeDepth's .Call boundary.

Host-side mirror of the reference's interface for ONE hot path (emission log-likelihood + HMM Viterbi):

    get_loglike_matrix(phi, expected, total, observed, mixture)   src/CNV_estimate.cpp:52-85
    C_hmm / viterbi_hmm(transitions, loglikelihood, positions, L)  src/hmm.cpp:18-167, R/tools.R:88-103
    ExomeDepth(test, reference, phi, expected, ...)                R/class_definition.R:82-191 (likelihood step)
    CallCNVs(x, chromosome, start, end, name, ...)                 R/class_definition.R:311-419
    TestCNV(x, chromosome, start, end, type)                       R/class_definition.R:243-256
    somatic_CNV_call(normal, tumor, prop_tumor, ...)               R/class_definition.R:442-461
    Cohort                                                         many samples × one bin set, device resident
and, widening along SURVEY.md §8f:
    betabin.fit(test, reference)                                   stand-in for aod::betabin (R/class_definition.R:118-119)
    refset.select_reference_set(test, references, bin_length, ...) R/optimize_reference_set.R:51-148

All numerics run in hand-written sm_100a CUDA kernels through the C ABI in include/exomedepth_b200.h.
There is no CPU fallback.
"""
from ._lib import EDB200Error, device_info, init, launch_count  # noqa: F401
from .api import C_hmm, CallCNVs, ExomeDepth, TestCNV, emission, get_loglike_matrix, lnbeta, somatic_CNV_call, viterbi_hmm  # noqa: F401
from .cohort import Cohort, pack_counts, pack_counts12, pack_counts12_numpy, pack_counts_numpy  # noqa: F401
from . import betabin, refset  # noqa: F401
