/* A plain-C caller that uses ONLY the reference's C ABI (raytracer.h: load_world, render,
 * move_camera_position) the way examples/c_raytracer.rs and GameView.swift do:
 *   c_caller <world.txt> <width> <height> <out.rgba> [dx dy dz]
 * loads the world text, optionally moves the camera once, renders one frame (16 spp, depth 8,
 * lib.rs:51) into a malloc'ed buffer and dumps the RGBA8 bytes.  Compiled and run by
 * tests/test_gpu_parity.py against libraytracer.so. */
#include "raytracer.h"

#include <stdio.h>
#include <stdlib.h>

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *source = (char *)malloc((size_t)n + 1);
    if (fread(source, 1, (size_t)n, f) != (size_t)n) return 3;
    source[n] = 0;
    fclose(f);

    size_t width = (size_t)atol(argv[2]), height = (size_t)atol(argv[3]);
    Rust_WorldHandle *handle = load_world(source);
    if (!handle) return 4;
    if (argc >= 8)
        handle->camera = move_camera_position(handle->camera, (float)atof(argv[5]), (float)atof(argv[6]), (float)atof(argv[7]));

    Rust_ColorU8 *pixels = (Rust_ColorU8 *)malloc(width * height * sizeof(Rust_ColorU8));
    Rust_CFramebuffer fb = { width, height, pixels };
    Rust_CFramebuffer out = render(fb, handle);
    if (out.pixels != pixels || out.width != width || out.height != height) return 5;

    f = fopen(argv[4], "wb");
    if (!f) return 6;
    fwrite(out.pixels, sizeof(Rust_ColorU8), width * height, f);
    fclose(f);
    return 0;
}
