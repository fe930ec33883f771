// apps/render_cli.cpp — the reference's `main()`s as one host program over the B200 core:
// builds a scene with the reference's own scene-description classes, renders it through
// the C ABI and writes a P3 PPM on stdout / timing on stderr exactly like
// rt_in_one_weekend/main.cpp:292-360 and accelerated-rt-cuda/final.cu:155-246.
//   render_cli [--scene weekend|next_week|perlin|earth|light|cornell|smoke|final|obj] [--obj file.obj]
//              [--image map.ppm] [--width W] [--height H] [--spp N] [--seed S] [--device D]
//              [--binary | --png]   (image on stdout; scene numbers 1-8 of rt_next_week/cuda/main.cu:402-459)
//              [--gpus N]  sample split over N devices, one host thread and one context per device,
//                          frames combined by rt_reduce (ncclReduce over NVLink);
//                          --host-combine sums the downloaded frames on the CPU instead (no NCCL needed)
//              [--instancing]  translate / rotate_y as instances of a two-level BVH instead of baked vertices
//   render_cli --scene-file FILE.scene [--width W --height H --spp N ...]   (include/rtx/scene_file.h)
#include <chrono>
#include <cstring>
#include <iostream>
#include <thread>

#include <unistd.h>

#include "scenes.h"
#include "scene_file.h"

int main(int argc, char **argv) {
  std::string scene = "weekend", obj, image, scene_path;
  int W = 0, H = 0, spp = 0, device = 0, gpus = 1;
  unsigned long long seed = 1984;
  bool have_seed = false;
  bool binary = false, png = false, host_combine = false, instancing = false;
  for (int i = 1; i < argc; i++) {
    auto is = [&](const char *f) { return !strcmp(argv[i], f) && i + 1 < argc; };
    if (is("--scene")) scene = argv[++i];
    else if (is("--obj")) obj = argv[++i];
    else if (is("--image")) image = argv[++i];
    else if (is("--width")) W = atoi(argv[++i]);
    else if (is("--height")) H = atoi(argv[++i]);
    else if (is("--spp")) spp = atoi(argv[++i]);
    else if (is("--seed")) { seed = strtoull(argv[++i], nullptr, 10); have_seed = true; }
    else if (is("--scene-file")) scene_path = argv[++i];
    else if (is("--device")) device = atoi(argv[++i]);
    else if (is("--gpus")) gpus = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--binary")) binary = true;
    else if (!strcmp(argv[i], "--png")) png = true;
    else if (!strcmp(argv[i], "--host-combine")) host_combine = true;
    else if (!strcmp(argv[i], "--instancing")) instancing = true;
    else { std::cerr << "unknown argument " << argv[i] << "\n"; return 2; }
  }
  // The image goes to stdout; libraries write there too (NCCL prints its version banner on fd 1): point fd 1 at
  // stderr while working and give stdout back just before the image is written.
  fflush(stdout);
  const int image_fd = dup(1);
  dup2(2, 1);
  try {
    render_options opt;
    scene_file sf;
    if (!scene_path.empty()) {
      sf = load_scene_file(scene_path);
      opt = sf.opt;
      if (!W) W = sf.width;
      if (!H) H = sf.height;
      if (!spp) spp = sf.spp;
      if (!have_seed) seed = sf.opt.seed;
      scene = "file";
    }
    if (!W) W = 1200;
    if (!H) H = 800;
    if (!spp) spp = 500;
    opt.device = device;
    opt.seed = seed;
    opt.instancing = instancing || opt.instancing; // --instancing or the scene file's `flags instancing`
    const double aspect = double(W) / H;
    hittable_list world;
    hittable *root = nullptr;
    std::vector<hittable *> d_list(22 * 22 + 1 + 3 + 1);
    if (scene == "file") {
      root = &sf.world;
    } else if (scene == "weekend") { // main.cpp:292-311
      world = random_scene();
      root = &world;
    } else if (scene == "next_week") { // main.cu:402-407,462-465
      root = next_week_random_scene(d_list.data());
      opt.profile = RT_PROFILE_NEXT_WEEK;
      opt.sky_gradient = false;
      opt.background = color(0.70, 0.80, 1.00);
    } else if (scene == "obj") { // obj_render.cu:716-724
      if (obj.empty()) { std::cerr << "--scene obj needs --obj file.obj\n"; return 2; }
      root = obj_model(obj);
      opt.profile = RT_PROFILE_NEXT_WEEK;
      opt.sky_gradient = false;
      opt.background = color(0, 0, 0);
      opt.t_min = 0.00001;
      opt.flags = RT_FLAG_FLIP_NORMALS | RT_FLAG_DEPTH_BACKGROUND;
    } else if (scene == "cornell") { // main.cu:436-443
      root = cornell_box();
      opt.profile = RT_PROFILE_NEXT_WEEK;
      opt.sky_gradient = false;
      opt.background = color(0, 0, 0);
    } else if (scene == "perlin" || scene == "earth" || scene == "light" || scene == "smoke" || scene == "final") {
      opt.profile = RT_PROFILE_NEXT_WEEK; // main.cu:417-459
      opt.sky_gradient = false;
      const bool sky = scene == "perlin" || scene == "earth";
      opt.background = sky ? color(0.70, 0.80, 1.00) : color(0, 0, 0);
      static std::vector<unsigned char> map;
      int mw = 512, mh = 256;
      if (scene == "earth" || scene == "final") {
        if (!image.empty()) {
          image_texture file(image.c_str());
          if (file.data.empty()) { std::cerr << "cannot read " << image << " (binary or ASCII PPM expected)\n"; return 2; }
          map = file.data; mw = file.width; mh = file.height;
        } else {
          map = procedural_earth(mw, mh);
        }
      }
      root = scene == "perlin" ? two_perlin_spheres()
             : scene == "earth" ? earth(map.data(), mw, mh)
             : scene == "light" ? simple_light()
             : scene == "smoke" ? cornell_smoke()
                                : rt_next_week_final_scene(map.data(), mw, mh);
    } else { std::cerr << "unknown scene " << scene << "\n"; return 2; }
    auto nw_cam = [&](point3 from, point3 at, double vfov) {
      return camera(from, at, vec3(0, 1, 0), vfov, aspect, 0.0, (from - at).length(), 0.0, 1.0, true);
    };
    camera cam = scene == "file" ? sf.make_camera(W, H)
                 : scene == "weekend"
                     ? camera(point3(13, 2, 3), point3(0, 0, 0), vec3(0, 1, 0), 20, aspect, 0.1, 10.0)
                 : (scene == "cornell" || scene == "smoke") ? nw_cam(point3(278, 278, -800), point3(278, 278, 0), 40)
                 : scene == "perlin" ? nw_cam(point3(13, 2, 3), point3(0, 0, 0), 20)
                 : scene == "earth" ? nw_cam(point3(13, 2, 3), point3(0, 0, 0), 40)
                 : scene == "light" ? nw_cam(point3(26, 3, 6), point3(0, 2, 0), 20)
                 : scene == "final" ? nw_cam(point3(478, 278, -600), point3(278, 278, 0), 40)
                 : scene == "next_week"
                     ? camera(point3(13, 2, 3), point3(0, 0, 0), vec3(0, 1, 0), 20, aspect, 0.05,
                              (point3(13, 2, 3) - point3(0, 0, 0)).length(), 0.0, 1.0, true)
                     : camera(point3(1, 3, 7), point3(0, 2, 0), vec3(0, 1, 0), 60, aspect, 0.0,
                              (point3(1, 3, 7) - point3(0, 2, 0)).length(), 0.0, 1.0, true);
    renderer r(opt);
    r.set_scene(*root, cam);
    std::cerr << "Rendering a " << W << "x" << H << " image with " << spp << " samples per pixel ("
              << r.flat.spheres.size() << " spheres, " << r.flat.triangles.size() << " triangles, " << r.flat.quads.size()
              << " rects, " << r.flat.media.size() << " media, " << r.flat.instances.size() << " instances of "
              << r.flat.groups.size() << " objects)\n";
    // further devices: the same scene, each its own context; device g renders the global sample
    // indices [g*spp/G, (g+1)*spp/G) (counter-based RNG: the image does not depend on G)
    if (gpus < 1 || device + gpus > rt_device_count()) throw std::runtime_error("--gpus: not that many CUDA devices");
    std::vector<std::unique_ptr<renderer>> more;
    for (int g = 1; g < gpus; g++) {
      render_options og = opt;
      og.device = device + g;
      more.emplace_back(new renderer(og));
      more.back()->set_scene(*root, cam);
    }
    auto share = [&](int g) { return std::make_pair((int)((long long)spp * g / gpus), (int)((long long)spp * (g + 1) / gpus)); };
    if (gpus > 1 && !host_combine) {
      std::vector<renderer *> all{&r};
      for (auto &m : more) all.push_back(m.get());
      renderer::group(all);
    }
    auto t0 = std::chrono::steady_clock::now();
    {
      // one host thread per device: render its share, then join the reduce (a single-process NCCL group needs the
      // ranks' calls issued concurrently)
      std::vector<std::thread> workers;
      std::vector<std::string> errors(gpus);
      auto job = [&](renderer &rr, int g) {
        try {
          rr.render(W, H, share(g).second - share(g).first, share(g).first);
          if (gpus > 1 && !host_combine) rr.reduce(0);
        } catch (const std::exception &e) { errors[g] = e.what(); }
      };
      for (int g = 1; g < gpus; g++) workers.emplace_back([&, g]() { job(*more[g - 1], g); });
      job(r, 0); // never throws: the workers are always joined
      for (auto &w : workers) w.join();
      for (auto &e : errors) if (!e.empty()) throw std::runtime_error(e);
    }
    if (gpus > 1 && host_combine) { // sum of the downloaded float4 frames (fixed device order)
      std::vector<float> sum = r.accum_download();
      for (auto &m : more) {
        std::vector<float> f = m->accum_download();
        for (size_t k = 0; k < sum.size(); k++) sum[k] += f[k];
      }
      r.accum_upload(W, H, sum);
    }
    image8 im = r.resolve();
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    rt_stats_t st = r.stats();
    for (auto &m : more) { rt_stats_t sg = m->stats(); st.segments += sg.segments; st.ms_render = std::max(st.ms_render, sg.ms_render); }
    std::cerr << "took " << secs << " seconds. (" << st.segments / 1e6 / (st.ms_render * 1e-3) << " Mpath-bounces/s, BVH "
              << st.n_nodes << " nodes built in " << st.ms_build << " ms)\n";
    fflush(stdout);
    std::cout.flush();
    dup2(image_fd, 1);
    close(image_fd);
    if (png) im.write_png(std::cout);
    else if (binary) im.write_ppm_binary(std::cout);
    else im.write_ppm(std::cout);
  } catch (const std::exception &e) {
    std::cerr << "error: " << e.what() << "\n";
    return 99; // the reference exits 99 on CUDA errors (final.cu:22)
  }
  return 0;
}
