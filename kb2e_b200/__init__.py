"""kb2e_b200 -- B200-native (sm_100a) implementation of the KB2E translation-embedding hot path.

The product is the C-ABI library ``kb2e_b200/lib/libkb2e_b200.so`` (include/kb2e_b200.h) and the
six reference-named executables in ``kb2e_b200/bin``; this package is the thin Python mirror of
that ABI used by the tests, the bench and the multi-GPU launcher.  There is no CPU fallback:
importing works anywhere, creating a ``Context`` needs a B200.
"""
from .api import (Context, Kb2eError, load_library, MODELS, TABLE_ENTITY, TABLE_RELATION, TABLE_WEIGHTS,  # noqa: F401
                  FLAG_RANK_EXACT_ONLY, FLAG_TRANSR_NO_QUIRK, FLAG_SAMPLER_RANDMAX, FLAG_DETERMINISTIC)
