{
  "targets": [
    {
      "target_name": "hzwitness_napi",
      "sources": ["hzwitness_napi.cc"],
      "include_dirs": ["../../include"],
      "libraries": ["<(module_root_dir)/../../circuits_b200/libhzwitness.so"],
      "ldflags": ["-Wl,-rpath,<(module_root_dir)/../../circuits_b200"],
      "cflags_cc": ["-std=c++17", "-O2"]
    }
  ]
}
