"""Golden vectors for the reference's block-filled ("SCREEN_SCALE" / progressive resolution) frames
(Raytracer.cpp:233-248, 330-341): its own renderArea, 16 column strips, preview shading (deterministic),
one overwrite frame. Run here (needs /root/reference and oracle/_ref); writes tests/golden/scaled.npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle_py import Oracle, Reference, REFERENCE_DIR  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")


def main():
    ref, orc = Reference(), Oracle()
    out = {}
    for scene in ("Scene1", "Scene3"):
        ref.load_scene(os.path.join(REFERENCE_DIR, "Scenes", scene + ".json"))
        for (w, h) in ((160, 120), (333, 77)):
            for screen_scale, scaler in ((0.5, 1.0), (0.5, 0.25), (1.0, 0.25), (0.3, 1.0)):
                ref.setup(w, h, 55, 2, True, orc.default_camera(55), screen_scale=screen_scale)
                ref.render_frames(1, rng_mode=0, start_frame=1, progressive_scaler=scaler)
                steps = int(np.ceil(1 / (np.float32(screen_scale) * np.float32(scaler))))
                out["%s_%dx%d_steps%d" % (scene, w, h, steps)] = ref.color_buffer()[..., :3].copy()
    np.savez_compressed(os.path.join(GOLD, "scaled.npz"), **out)
    print(sorted(out))


if __name__ == "__main__":
    main()
