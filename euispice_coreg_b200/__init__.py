"""B200-native lag-grid pointing search with the API of adolliou/euispice_coreg's hot path.

    from euispice_coreg_b200.hdrshift import Alignment, AlignmentResults
    from euispice_coreg_b200.synras import SPICEComposedMapBuilder

See DESIGN.md for the scope and INTEGRATION.md for the C ABI.
"""
__version__ = "0.1.0"
