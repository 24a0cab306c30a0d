"""The only rendered golden the reference ships: docs/src/assets/shadows-sppm-1024x1024_mio.png, the published output of
docs/code/spheres.jl (README.md:9-11).  tests/golden/shadows_reference_128.npy is that image box-filtered to 128x128
(scripts/make_golden_shadows.py).  The oracle's SPPM render of the same scene at 128x128 - camera algebra with its quirks,
sphere / triangle intersection, Matte / Mirror / Glass shading, point light, SPPM passes, film tone path and row flip -
must reproduce it: this pins the restatement end to end against an output of the reference itself.  The golden is 8-bit
with an unknown iteration count, so the comparison is structural: mean level, luminance correlation, and an 8x8 map of
block means (observed: mean 0.385 vs 0.380, correlation 0.93, largest block difference 0.07)."""
import os

import numpy as np
import pytest

import oracle_lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shadows_reference_128.npy")


def _lum(a):
    return 0.2126 * a[..., 0] + 0.7152 * a[..., 1] + 0.0722 * a[..., 2]


def _blocks(a):
    return a.mean(axis=2).reshape(8, 16, 8, 16).mean(axis=(1, 3))


def _compare(img, gold):
    corr = float(np.corrcoef(_lum(img).ravel(), _lum(gold).ravel())[0, 1])
    mean_diff = abs(float(img.mean()) - float(gold.mean()))
    block_diff = float(np.abs(_blocks(img) - _blocks(gold)).max())
    return corr, mean_diff, block_diff


def test_oracle_reproduces_the_published_shadows_image(T):
    gold = np.load(GOLDEN).astype(np.float32) / 255.0
    scene, camera, kw = T.scenes.shadows(resolution=128)
    osc = oracle_lib.OracleScene(scene.flatten())
    cam, fd = camera.pod(), camera.film.desc()
    rgb = np.zeros(camera.film.pixels.shape[:2] + (3,), np.float32)
    osc.render_sppm(cam, fd, kw["initial_search_radius"], kw["max_depth"], 20, 200_000, 1, rgb)
    camera.film.set_image(rgb)
    img = camera.film.to_rgb()[::-1]                     # film.save flips the rows (film.jl:221)
    corr, mean_diff, block_diff = _compare(img, gold)
    assert corr > 0.9 and mean_diff < 0.02 and block_diff < 0.12, (corr, mean_diff, block_diff)
    # the back wall's horizontal falloff, top block row, to two decimals of the published image
    assert np.abs(_blocks(img)[0] - _blocks(gold)[0]).max() < 0.02


@pytest.mark.gpu
def test_gpu_reproduces_the_published_shadows_image(T, ctx):
    gold = np.load(GOLDEN).astype(np.float32) / 255.0
    scene, camera, kw = T.scenes.shadows(resolution=128)
    integrator = T.SPPMIntegrator(camera, kw["initial_search_radius"], kw["max_depth"], 20, 200_000, write_frequency=20, context=ctx)
    integrator(scene)
    img = camera.film.to_rgb()[::-1]
    corr, mean_diff, block_diff = _compare(img, gold)
    assert corr > 0.9 and mean_diff < 0.02 and block_diff < 0.12, (corr, mean_diff, block_diff)
    assert np.abs(_blocks(img)[0] - _blocks(gold)[0]).max() < 0.02
