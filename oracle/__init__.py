"""TEST INFRASTRUCTURE ONLY: CPU oracles for the single histogram filter.

Importable only from tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs. The product package
``superterrainplus_b200`` never imports this package.
"""
from .pyoracle import (  # noqa: F401
    BIN_DTYPE,
    BIOME_PROPERTY_DTYPE,
    OracleError,
    build,
    closed_form,
    have_height_reference,
    have_reference,
    heightfield_port,
    heightfield_reference,
    reference_pinned_active,
    run_port,
    run_reference,
    ReferenceSession,
)
