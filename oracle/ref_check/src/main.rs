// ref_check — the reference's own CLI render (raytracer/src/main.rs:48-103) with the world file, the output path
// and samples / depth taken from argv instead of the hard-coded /Users/... path of parser::parse_world()
// (parser.rs:47-52).  Every call below is the unmodified crate's public API, in main.rs's order:
//   parser::parse_input -> World::new(spheres, vec![mesh]) -> Camera::new_look_at((0,0,0), (0,0,-1), Y, PI/2, 1.77778)
//   -> image 400 x (400 / aspect) -> ray_trace -> write_image (ASCII P3).
// The oracle's SERIAL mode (oracle/rt_oracle.c, one xorshift32 stream seeded 2547549 for the whole frame) must
// reproduce the PPM this writes byte for byte: `python check.py --rust-ppm <file>`.
use raytracer::camera::{Camera, Radians};
use raytracer::common::{ray_trace, Options, World};
use raytracer::image::{write_image, Framebuffer};
use raytracer::maths::{Vec3, Y_AXIS};
use raytracer::parser;

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() < 3 {
        eprintln!("usage: ref_check <world.txt> <out.ppm> [samples_per_pixel=50] [max_ray_bounces=8] [image_width=400]");
        std::process::exit(2);
    }
    let samples: i32 = args.get(3).map(|s| s.parse().unwrap()).unwrap_or(50); // main.rs:23
    let depth: i32 = args.get(4).map(|s| s.parse().unwrap()).unwrap_or(8); // main.rs:24
    let image_width: usize = args.get(5).map(|s| s.parse().unwrap()).unwrap_or(400); // main.rs:91

    let text = std::fs::read_to_string(&args[1]).expect("cannot read the world file");
    let (_camera, spheres, mesh) = parser::parse_input(&text).expect("parse error"); // main.rs:57
    let world = World::new(spheres, vec![mesh]); // main.rs:58-59

    let camera = Camera::new_look_at(
        Vec3::new(0.0, 0.0, 0.0),
        Vec3::new(0.0, 0.0, -1.0),
        Y_AXIS.into(),
        Radians(std::f32::consts::PI / 2.0),
        1.77778,
    ); // main.rs:86-88

    let aspect_ratio = camera.aspect_ratio(); // main.rs:90
    let image_height = (image_width as f32 / aspect_ratio) as usize; // main.rs:92

    let mut options = Options::new(samples, depth, None, true); // main.rs:51 without the stderr logger
    let framebuffer = Framebuffer::new(image_width, image_height);
    let framebuffer = ray_trace(&world, &camera, framebuffer, &mut options); // main.rs:95
    write_image(&framebuffer, Some(&args[2])).expect("cannot write the image"); // main.rs:99
    eprintln!("{}x{} {} spp depth {} -> {}", image_width, image_height, samples, depth, args[2]);
}
