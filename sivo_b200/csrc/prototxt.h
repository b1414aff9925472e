// Topology loader for the Bayesian SegNet prototxts -- replaces `new caffe::Net<float>(model_file, TEST)`
// (src/bayesian_segnet/bayesian_segnet.cpp:59-60) for the nine layer types those files use.
#pragma once
#include <string>
#include <vector>

namespace sivo {

enum class LayerType { Convolution, ReLU, BN, LRN, Pooling, Upsample, Dropout, Softmax };

struct LayerSpec {
  std::string name;
  LayerType type;
  std::vector<std::string> bottoms, tops;
  // Convolution
  int num_output = 0, kernel = 0, pad = 0;
  bool bias_term = true;
  // LRN
  int local_size = 5;
  float alpha = 1.f, beta = 0.75f, k = 1.f;
  // Dropout
  float dropout_ratio = 0.5f;
  bool sample_weights_test = false;
  // ReLU
  float negative_slope = 0.f;
};

struct NetSpec {
  std::string name, input_name;
  int dims[4] = {0, 0, 0, 0};  // N (= T, 0 if left blank), C, H, W
  std::vector<LayerSpec> layers;
};

NetSpec parse_prototxt_file(const std::string& path);
NetSpec parse_prototxt_text(const std::string& text);

}  // namespace sivo
