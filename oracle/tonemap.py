"""tonemap.py — CPU ORACLE, TEST INFRASTRUCTURE (not product code): restatement of the reference's display pass,
assets/shaders/tone-map.frag as run by rfw::system::render_frame(camera, status, toneMap=true) (system/src/rfw/system.cpp:694-713,
params = (camera.contrast, camera.brightness, 0, 0)), followed by the UNORM8 conversion of an 8-bit target.

PARITY PINNING: the reference is a GLSL fragment shader and holds no vectors for it; it cannot be executed here (no GL).
`tone_map_spec` evaluates the shader's formulas in float64 as the specification; `tone_map` evaluates them in float32, one
rounding per operation in the shader's order, which is what rfwb200's k_tone_map does — bytes must match `tone_map` exactly
and `tone_map_spec` within one code value.  parity unpinned (restated from source, cited above)."""
import numpy as np

# mat3(vec3 c0, vec3 c1, vec3 c2) of the shader: columns
ACES_IN = np.array([[0.59719, 0.07600, 0.02840], [0.35458, 0.90834, 0.13383], [0.04823, 0.01566, 0.83777]])
ACES_OUT = np.array([[1.60475, -0.10208, -0.00327], [-0.53108, 1.10813, -0.07276], [-0.07367, -0.00605, 1.07602]])


def _fit(v, dt):
    c = lambda x: dt(x)  # noqa: E731
    a = v * (v + c(0.0245786)) - c(0.000090537)
    b = v * (c(0.983729) * v + c(0.4329510)) + c(0.238081)
    return a / b


def _mat(cols, rgb, dt):
    cols = cols.astype(dt)
    # M * v = c0 * v.x + c1 * v.y + c2 * v.z, summed left to right
    return np.stack([(cols[0, k] * rgb[..., 0] + cols[1, k] * rgb[..., 1]) + cols[2, k] * rgb[..., 2] for k in range(3)], -1)


def _tone(rgba, contrast, brightness, dt):
    rgba = np.asarray(rgba).astype(dt)
    rgb = np.maximum(dt(0), ((rgba[..., :3] - dt(0.5) * dt(contrast)) + dt(0.5)) + dt(brightness))
    out = _mat(ACES_OUT, _fit(_mat(ACES_IN, rgb, dt), dt), dt)
    return np.concatenate([np.clip(out, dt(0), dt(1)), np.clip(rgba[..., 3:4], dt(0), dt(1))], -1)


def tone_map(rgba, contrast=0.0, brightness=0.0):
    """float32, operation for operation -> uint8 (…, 4)"""
    with np.errstate(invalid="ignore", divide="ignore"):
        t = _tone(rgba, contrast, brightness, np.float32)
        return (np.nan_to_num(t, nan=0.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)


def tone_map_spec(rgba, contrast=0.0, brightness=0.0):
    """float64 evaluation of the shader -> float values in [0, 255] before rounding"""
    return _tone(rgba, contrast, brightness, np.float64) * 255.0
