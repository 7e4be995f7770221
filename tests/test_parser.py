"""The product's world-text parser (load_world, through the C ABI) against the oracle's
restatement of parser.rs:336-382: same accept/reject decision, same primitives, same camera."""
import numpy as np
import pytest

VALID = [
    "camera origin 0.0 0.0 0.0 aspect 1.77778;",
    "camera origin 0.0 0.0 0.0 aspect 1.77778;\n\n",
    "// a comment\ncamera origin 1.5 -2.25 3.0 aspect 2.0;\nmaterial A : Diffuse color 0.1 0.2 0.3;\n"
    "sphere center 0.0 0.0 -1.0 radius 0.5 material A;\n",
    "camera origin 0 0 0 aspect 1.0;material M:Metal color 1 .5 0. fuzz 0.25;sphere center -1 -.5 -2. radius 1 material M;",
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial G : Dielectric ir 1.5;\n// c1\n// c2\n"
    "material G : Diffuse color 1.0 0.0 0.0;\nsphere center 0.0 0.0 -1.0 radius 0.5 material G;\n",   # redefinition wins
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial R : Diffuse color 1.0 0.0 0.0;\n"
    "triangle v0 -0.1 -0.1 -0.5 v1 0.1 -0.1 -0.5 v2 -0.1 0.1 -0.5 material R;\n",
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial R : Diffuse color 1.0 0.0 0.0;\n"
    "sphere center 0.0 0.0 -1.0 radius 0.5 material R;  \n",                                   # Unicode whitespace
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial R : Diffuse color 007.50 00.0 1.;\n",
    "camera\torigin\n0.0\r\n0.0 0.0 aspect 1.0 ;",
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial  : Diffuse color 1.0 0.0 0.0;\nsphere center 0.0 0.0 -1.0 radius 0.5 material ;\n",  # empty name
]

# parser.rs:59-62: material names are runs of char::is_alphanumeric() (UNICODE Alphabetic / Numeric) or '_'
VALID += [
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial RÖD_färg : Diffuse color 1.0 0.0 0.0;\n"
    "sphere center 0.0 0.0 -1.0 radius 0.5 material RÖD_färg;\n",                                  # Latin-1 letters
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial 金属٣Ⅻ : Metal color 0.5 0.5 0.5 fuzz 0.1;\n"
    "sphere center 0.0 0.0 -1.0 radius 0.5 material 金属٣Ⅻ;\n",                                     # CJK, Arabic-Indic digit (Nd), Roman numeral (Nl)
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial 𝔊𐐷 : Dielectric ir 1.5;\n"
    "sphere center 0.0 0.0 -1.0 radius 0.5 material 𝔊𐐷;\n",                                        # 4-byte scalars (SMP letters)
]

INVALID = [
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial A→B : Diffuse color 1.0 0.0 0.0;\n",           # U+2192 is not alphanumeric
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial A😀 : Diffuse color 1.0 0.0 0.0;\n",           # emoji: So
    "",
    " camera origin 0.0 0.0 0.0 aspect 1.0;",                          # leading whitespace: MissingCamera
    "// c\n\ncamera origin 0.0 0.0 0.0 aspect 1.0;",                   # blank line after comment: MissingCamera
    "// unterminated comment",
    "camera origin 0.0 0.0 aspect 1.0;",
    "camera origin 0.0 0.0 0.0 aspect 1.0",                            # missing ';'
    "camera origin 0.0 0.0 0.0 aspect 1;",                             # "1;" is < 3 bytes of input left (parser.rs:112)
    "camera origin 0.0 0.0 0.0 aspect +1.0;",
    "camera origin 0.0 0.0 0.0 aspect 1e0;",
    "camera origin 0.0 0.0 0.0 aspect 1.0.0;",
    "camera origin 0.0 0.0 0.0 aspect -.;;",
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nsphere center 0.0 0.0 -1.0 radius 0.5 material NOPE;\n",
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial A : Plastic color 1.0 1.0 1.0;\n",
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial A : Emission color 1.0 1.0 1.0;\n",   # parser.rs:171-174: no Emission
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial A : Diffuse color 1.0 1.0 1.0;\n// c\n\nmaterial B : Diffuse color 1.0 1.0 1.0;\n",
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial A : Diffuse color 1.0 1.0 1.0;\n"
    "triangle v0 0.0 0.0 0.0 v1 1.0 0.0 0.0 v2 0.0 1.0 0.0 material A;\nsphere center 0.0 0.0 -1.0 radius 0.5 material A;\n",  # order
    "camera origin 0.0 0.0 0.0 aspect 1.0;\ntrailing garbage",
    "camera origin 0.0 0.0 0.0 aspect 1.0;\nmaterial A : Metal color 1.0 1.0 1.0;\n",       # fuzz missing
]


@pytest.mark.parametrize("i", range(len(VALID)))
def test_valid_worlds_parse_identically(rt, ob, i):
    text = VALID[i]
    cam, w = ob.parse_input(text)
    h = rt.load_world(text)
    assert h.n_spheres == w.n_spheres and h.n_triangles == w.n_triangles
    assert np.array_equal(h.camera_floats(), cam.floats())
    for k, s in enumerate(w.spheres()):
        want = [s.center.x, s.center.y, s.center.z, s.radius, s.material.type, s.material.r, s.material.g,
                s.material.b, s.material.param]
        assert np.array_equal(h.sphere(k), np.array(want, dtype=np.float32))
    for k, t in enumerate(w.triangles()):
        want = [t.v0.x, t.v0.y, t.v0.z, t.v1.x, t.v1.y, t.v1.z, t.v2.x, t.v2.y, t.v2.z, t.normal.x, t.normal.y,
                t.normal.z, t.material.type, t.material.r, t.material.g, t.material.b, t.material.param, 0.0]
        assert np.array_equal(h.triangle(k), np.array(want, dtype=np.float32), equal_nan=True)


@pytest.mark.parametrize("i", range(len(INVALID)))
def test_invalid_worlds_are_rejected_by_both(rt, ob, i):
    text = INVALID[i]
    with pytest.raises(ob.ParseError):
        ob.parse_input(text)
    with pytest.raises(rt.ParseError):
        rt.load_world(text)
    assert rt.last_error().startswith("load_world: ParseError")


def test_invalid_utf8_is_rejected(rt, ob):
    bad = b"camera origin 0.0 0.0 0.0 aspect 1.0;\n// \xff\xfe\n"
    with pytest.raises(ob.ParseError):
        ob.parse_input(bad)
    with pytest.raises(rt.ParseError):
        rt.load_world(bad)


def test_generated_scenes_parse(rt, ob, scenes):
    for text, ns, nt in ((scenes.default_world(), 8, 0), (scenes.example_world(), 8, 2),
                         (scenes.c3_world(), 1000, 0), (scenes.c5_world(), 8000, 2000)):
        h = rt.load_world(text)
        assert (h.n_spheres, h.n_triangles) == (ns, nt)
        cam, w = ob.parse_input(text)
        assert (w.n_spheres, w.n_triangles) == (ns, nt)
        k = ns - 1
        s = w.spheres()[k]
        assert np.array_equal(h.sphere(k)[:4], np.array([s.center.x, s.center.y, s.center.z, s.radius], np.float32))


def test_default_world_matches_reference_scene(rt, scenes):
    """list order = hit-test order."""
    h = rt.load_world(scenes.default_world())
    want = [((0, -100.5, -1), 100.0, 0, (0.8, 0.8, 0.0), 0.0), ((0, 0, -1), 0.5, 0, (0.7, 0.3, 0.3), 0.0),
            ((-1, 0, -1), 0.5, 1, (0.8, 0.8, 0.8), 0.3), ((1, 0, -1), 0.5, 2, (1, 1, 1), 1.5),
            ((0, 1, -2), 0.5, 1, (0.9, 0.9, 0.9), 0.0), ((-3, 2, -3), 0.5, 0, (1, 0, 0), 0.0),
            ((0, 2, -3), 0.5, 0, (0, 1, 0), 0.0), ((3, 2, -3), 0.5, 0, (0, 0, 1), 0.0)]
    assert h.n_spheres == 8
    for k, (c, r, t, col, p) in enumerate(want):
        assert np.array_equal(h.sphere(k), np.array([*c, r, t, *col, p], dtype=np.float32))


def test_emission_extension_and_world_to_text_round_trip(rt, scenes):
    """`Emission color r g b` (MaterialType::Emission exists, materials.rs:11, but the
    reference grammar cannot express it) and the inverse of load_world with exact decimal floats."""
    src = scenes.example_world().replace("material GLASS : Dielectric ir 1.5;",
                                         "material GLASS : Dielectric ir 1.5;\nmaterial LAMP : Emission color 4.0 3.5 0.25;")
    src = src.replace("radius 0.5 material MIRROR", "radius 0.5 material LAMP")
    with pytest.raises(rt.ParseError):
        rt.load_world(src)                               # load_world stays the reference's grammar
    h = rt.load_world(src, rt.PARSE_EMISSION)
    assert any(h.sphere(i)[4] == rt.EMISSION and list(h.sphere(i)[5:8]) == [4.0, 3.5, 0.25] for i in range(h.n_spheres))
    # awkward floats: tiny, huge, negative zero, non-terminating binary fractions
    g = rt.world_new((0.1, -0.0, 1e-7), 1.77778)
    g.add_sphere((1e-20, 123456.789, -3.3333333), 0.1, rt.METAL, (0.123456789, 1e-9, 0.999999), 0.3)
    g.add_sphere((0, 0, -1), 7.5e8, rt.EMISSION, (2, 3, 4))
    g.add_triangle((0.1, 0.2, 0.3), (1.1, 0.2, 0.3), (0.1, 1.2, 0.35), rt.DIELECTRIC, (1, 1, 1), 1.5)
    for w in (h, g):
        text = w.to_text()
        assert "e" not in text.split("Emission")[0].replace("sphere", "").replace("center", "").replace("triangle", "")\
            .replace("material", "").replace("Dielectric", "").replace("Metal", "").replace("Diffuse", "").replace("aspect", "")\
            .replace("camera", "")                                # no exponents anywhere (parse_float has none)
        w2 = rt.load_world(text, rt.PARSE_EMISSION)
        assert (w2.n_spheres, w2.n_triangles) == (w.n_spheres, w.n_triangles)
        for i in range(w.n_spheres):
            assert w2.sphere(i).tobytes() == w.sphere(i).tobytes()
        for j in range(w.n_triangles):
            assert w2.triangle(j).tobytes() == w.triangle(j).tobytes()
        assert w2.camera_floats().tobytes() == w.camera_floats().tobytes()
