"""A deliberately awkward OBJ + scene (quads, a concave pentagon, negative indices, `o` and `g`
tags, empty groups, CRLF, exponents, more groups than materials, rotation + non-uniform scale)
shared by make_golden.py (reference side) and tests/test_frontend.py (product side)."""
import json
import os

OBJ = "\r\n".join([
    "# objedge", "mtllib nothing.mtl", "v 0 0 0", "v 1 0 0", "v 1 1 0", "v 0 1 0", "v 0.5 0.5 1.0e0", "v 2 0 0.25", "v 2 1.5 -0.5",
    "v 3.25e-1 2 1", "v -1 -1 -1", "v 1.5 0.2 0.1", "v 2.0 2.0 0.0", "v 0.9 0.9 0.0", "v 0.0 2.0 0.0", "v -0.5 1.0 0.0",
    "vt 0 0", "vt 1 0", "vt 1 1", "vt 0 1", "vt 0.5 0.5",
    "vn 0 0 1", "vn 0 1 0", "vn 1 0 0", "vn 0.57735 0.57735 0.57735", "vn 0 0 -2",
    "g first", "usemtl whatever", "f 1/1/1 2/2/1 3/3/1 4/4/1", "f 1/1/2 2/2/2 5/5/4",
    "g emptygroup", "g second third", "f -13/1/3 -12/2/3 -10/3/3", "f 6/1/1 7/2/2 8/3/3 5/4/4",
    "o objecttag", "f 2/2/5 6/1/5 7/3/5", "# concave pentagon", "f 1/1/1 10/2/1 11/3/1 12/5/1 13/4/1",
    "g fourth", "f 4/4/1 14/1/1 13/2/1", "g fifth", "f 9/1/2 1/2/2 14/3/2", ""])

SCENE = {
    "Background": {"Name": "ptbsky64", "Path": "res\\texture\\", "Format": "bmp"},
    "Material": [{"Name": "edge_tex", "Diffuse": "0.5 0.6 0.7", "Emission": "0.0 0.0 0.0", "Specular": "0.2 0.2 0.2", "Transparent": "false",
                  "Roughness": "1.7", "RefractionIndex": "1.4", "ExtinctionCoef": "0.0", "AbsorptionCoef": "0.1 0.2 0.3",
                  "ReducedScatteringCoef": "0.0 0.0 0.0", "DiffuseTextureId": "-1"},
                 {"Name": "glass", "Diffuse": "0.9 0.9 0.9", "Emission": "0.0 0.0 0.0", "Specular": "0.1 0.1 0.1", "Transparent": "true",
                  "Roughness": "-3", "RefractionIndex": "1.6", "ExtinctionCoef": "0.0", "AbsorptionCoef": "0.0 0.0 0.0",
                  "ReducedScatteringCoef": "1.0 2.0 3.0"}],
    "Sphere": [{"Material": "glass", "Center": "0.5 0.25 -2.0", "Radius": "-1.0"}, {"Material": "zinc", "Center": "1 2 3", "Radius": "0.5"}],
    "Mesh": [{"Material": ["edge_tex", "gold", "wall_red"], "Path": "res\\obj\\objedge.obj", "Position": "0.5 -0.25 1.0", "Scale": "1.5 0.5 2.0",
              "Rotate": "30.0 -45.0 10.0"},
             {"Material": ["light"], "Path": "res\\obj\\objedge.obj", "Position": "0.0 3.0 0.0", "Scale": "1.0 1.0 -1.0", "Rotate": "0.0 0.0 90.0"}],
}


def write(root):
    """Writes the OBJ and scene into an existing workload root (needs res/texture/ptbsky64)."""
    os.makedirs(os.path.join(root, "res", "obj"), exist_ok=True)
    os.makedirs(os.path.join(root, "res", "scene"), exist_ok=True)
    with open(os.path.join(root, "res", "obj", "objedge.obj"), "w", newline="") as f:
        f.write(OBJ)
    path = os.path.join(root, "res", "scene", "objedge.json")
    with open(path, "w") as f:
        json.dump(SCENE, f, indent=1)
    return path
