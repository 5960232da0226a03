"""Host-side pieces next to the hot path: PNG capture and the real-time camera /
frame state machine of index.tsx:61-283.  No GPU."""
import math

import numpy as np

from raymarching_engine_b200 import png, viewer


def test_png_round_trip_and_orientation():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (5, 7, 4), dtype=np.uint8)
    data = png.encode_png(img)
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    np.testing.assert_array_equal(png.decode_png(data), img)
    # first scanline in the file is the TOP of the picture = the last GL row
    import struct, zlib
    pos = 8
    idat = b""
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        if tag == b"IDAT":
            idat += data[pos + 8:pos + 8 + n]
        pos += 12 + n
    raw = zlib.decompress(idat)
    assert raw[0] == 0 and raw[1:1 + 7 * 4] == img[-1].tobytes()
    assert png.to_data_url(img).startswith("data:image/png;base64,iVBOR")


def test_mat4_rotate_matches_closed_form():
    m = viewer.mat4_rotate(viewer.mat4_identity(), 0.3, (0.0, 1.0, 0.0))
    c, s = math.cos(0.3), math.sin(0.3)
    want = np.array([c, 0, -s, 0, 0, 1, 0, 0, s, 0, c, 0, 0, 0, 0, 1], np.float32)
    np.testing.assert_allclose(m, want, atol=1e-7)
    # composition order: mouse x then mouse y post-multiplies (yaw, then pitch about the new x axis)
    m2 = viewer.mat4_rotate(m, 0.2, (1.0, 0.0, 0.0))
    fwd = viewer.vec3_transform_mat4((0.0, 0.0, 1.0), m2)
    np.testing.assert_allclose(fwd, (math.sin(0.3) * math.cos(0.2), -math.sin(0.2), math.cos(0.3) * math.cos(0.2)), atol=1e-6)
    assert viewer.mat4_rotate(m, 1.0, (0.0, 0.0, 0.0)).tolist() == m.tolist()      # degenerate axis: unchanged


def test_new_frame_rule_is_one_loop_late():
    c = viewer.RealtimeController(camera_speed=0.5)
    seq = []
    c.key("w", True)
    for i in range(3):
        seq.append(c.begin_loop()); c.end_loop()
    c.key("w", False)
    for i in range(3):
        seq.append(c.begin_loop()); c.end_loop()
    # (frameid of the loop's job, samples handed to the presenter).  Every loop that FOLLOWS a moving loop bumps frameid
    # and resets the counter - but its own job was built before the bump (index.tsx:121-182 vs :221-231), so it still
    # draws into the previous frame's buffers; the fresh buffers are first used one loop later
    assert seq == [(0, 1), (0, 1), (1, 1), (2, 1), (3, 2), (3, 3)]
    assert c.frameid == 3
    assert c.viewer_position == [0.0, 0.0, 1.5]            # three loops of +z at speed 0.5, identity rotation
    c.requesting_new_frame = True
    assert c.begin_loop() == (3, 1) and c.frameid == 4
    c.end_loop()
    assert c.begin_loop() == (4, 2)


def test_mouse_look_keeps_switching_for_five_loops_and_moves_along_view():
    c = viewer.RealtimeController(camera_speed=1.0)
    c.mouse_move(100.0, 0.0)                                  # yaw by 0.4 rad
    ids = []
    for _ in range(8):
        ids.append(c.begin_loop()[0]); c.end_loop()
    assert ids == [0, 0, 1, 2, 3, 4, 5, 5]                    # mouseHasMoved counts 5,4,3,2,1 -> five switching loops, one loop late
    c.key("w", True)
    c.begin_loop(); c.end_loop()
    np.testing.assert_allclose(c.viewer_position, [math.sin(0.4), 0.0, math.cos(0.4)], atol=1e-6)
    c.pointer_locked = False
    c.mouse_move(50.0, 50.0)                                  # ignored without pointer lock
    np.testing.assert_allclose(viewer.vec3_transform_mat4((0, 0, 1), c.camera_rotation), [math.sin(0.4), 0.0, math.cos(0.4)], atol=1e-6)
