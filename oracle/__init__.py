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
* `oracle.wcs_car`     -- the plate-carree (-CAR) chain of wcslib (celset, sphx2s / sphs2x, carx2s / cars2x) behind
                          `Alignment.align_using_initial_carrington` (`hdrshift/alignment.py:344-399`).
* `oracle.pxlshift`    -- `pxlshift/alignment_pixels.py:14-156`, `pxlshift/c_correlate.py:41-62`,
                          `utils/matrix_transform.py:77-106`.

PARITY PIN STATUS
-----------------
astropy / sunpy / matplotlib are absent from this image (no network; the reference's build backend poetry-core is
absent too), so `pip install /root/reference` fails and `import euispice_coreg.hdrshift` dies on `import astropy`.
The reference's own code nevertheless RUNS here behind import stand-ins (`tests/golden/_ref_standins.py`: inert modules
for matplotlib & co., a small `astropy.units.Quantity`, the product's pure-Python FITS reader as `astropy.io.fits`, the
standard library's shared memory as `multiprocess.shared_memory`, and `astropy.wcs.WCS` answered by `oracle.wcs_tan` /
`oracle.wcs_car`). The generators under `tests/golden/` drive it and commit its outputs as fixtures.

* pinned:   `oracle.hpc`, `oracle.carrington` and the engine semantics they restate (lag enumeration and cube axis
            order, `_shift_header`, PCi_j rebuild, thresholds, `fov_limits`, common grid, float32 rounding, masks,
            Pearson) against the cubes of the reference's own `Alignment(parallelism=True)` for six cases --
            helioprojective x4, Carrington "fa", initial Carrington -- bit for bit
            (`tests/golden/make_alignment_golden.py` -> `alignment_golden.npz`, `tests/test_reference_golden.py`).
            The Carrington "fa" chain (`utils/rectify.py`, NumPy dtype trail included) involves no WCS: fully pinned.
            `oracle.synras` (+ the SPICE search through `oracle.hpc`) against the reference's own
            `SPICEComposedMapBuilder.process` and `AlignmentSpice.align_using_helioprojective`: synthetic raster, composed
            header and three cubes bit for bit (`make_spice_golden.py` -> `spice_golden.npz`,
            `tests/test_reference_golden_spice.py`).
            `oracle.pxlshift` against the reference's own `AlignmentPixels.find_best_parameters`, bit for bit
            (`make_pxlshift_golden.py` -> `pxlshift_golden.npz`, `tests/test_pxlshift.py`).
            `oracle.pearson` against the reference's numba `c_correlate` (`make_pearson_golden.py`);
            `oracle.resample` against scipy's `map_coordinates` (the very function the reference calls) at test time;
            the host `AlignmentResults` against the reference's printed 11x6 golden cube
            (`hdrshift/test/test_AlignmentResults.py:35-126, 172-173`).
* UNPINNED: the arithmetic INSIDE `astropy.wcs.WCS` (wcslib) and `astropy.units`: `oracle.wcs_tan`, `oracle.wcs_car`
            and the 4-axis SPICE header handling are restated from FITS-WCS Paper II / wcslib's documented
            linp2x / tanx2s / carx2s / sphx2s / celset algorithms and cross-checked only against independent
            closed-form derivations (gnomonic formulas; sphere-rotation form of the CAR chain). The goldens above use
            these restatements as the reference's WCS, so they say nothing about wcslib itself. `oracle.rice`
            (cfitsio) is unpinned except for the dither sequence's documented check value.
"""
