"""CPU oracle for the lag-grid pointing search -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this package.
Nothing under `euispice_coreg_b200/` imports it; the product path raises when its CUDA library is
missing rather than falling back to anything here.

What it restates (file:line into adolliou/euispice_coreg v0.4.0, `parallelism=True` semantics):

* `oracle.wcs_tan`     -- the FITS-WCS TAN pixel<->world chain the reference obtains from
                          `astropy.wcs` (wcslib 8.x bundled in astropy 7.2.0, `poetry.lock:190-191`), call
                          sites `utils/Util.py:283-312`, `hdrshift/alignment.py:1038-1069`.
* `oracle.resample`    -- `scipy.ndimage.map_coordinates(order=k, mode='constant', prefilter=False)`
                          as called by `utils/Util.py:82-104` / `utils/rectify.py:22-56`; the real scipy
                          function is used, plus a numpy restatement checked bit-for-bit against it.
* `oracle.pearson`     -- `hdrshift/c_correlate.py:39-72` (two-pass, sequential-sum Pearson).
* `oracle.hpc`         -- `hdrshift/alignment.py:401-468, 509-549, 580-887, 987-1069` (helioprojective search).
* `oracle.carrington`  -- `utils/rectify.py:282-423, 865-888` + `hdrshift/alignment.py:889-901`.
* `oracle.synras`      -- `synras/map_builder.py:87-131`.

PARITY PIN STATUS
-----------------
The reference cannot be imported in this image (no astropy / sunpy / matplotlib, no network, and its
build backend poetry-core is absent), and it holds no offline golden vector for an image->correlation
computation (all of its integration tests download FITS files).

* pinned:   `oracle.pearson` against the reference's own numba `c_correlate`, loaded by file path in the
            build container (`tests/golden/make_pearson_golden.py` -> `tests/golden/pearson_golden.npz`);
            `oracle.resample` against scipy's `map_coordinates` (the very function the reference calls) at
            test time; the host `AlignmentResults` against the reference's printed 11x6 golden cube
            (`hdrshift/test/test_AlignmentResults.py:35-126, 172-173`).
* UNPINNED: `oracle.wcs_tan` (wcslib is absent; restated from FITS-WCS Paper II / wcslib's documented
            sphx2s/sphs2x/tanx2s/tans2x/linp2x algorithm and cross-checked only against an independent
            closed-form gnomonic derivation) and therefore the end-to-end cubes of `oracle.hpc`.
            "parity unpinned" applies to that boundary.
"""
