// Binary `.caffemodel` reader -- replaces `Net::CopyTrainedLayersFrom(weights_file)`
// (src/bayesian_segnet/bayesian_segnet.cpp:61; caffe/src/caffe/net.cpp:788-803,750-785).
// Hand-parses the protobuf wire format (no protoc in the image): NetParameter.layer = 100,
// LayerParameter{name = 1, type = 2, blobs = 7}, BlobProto{shape = 7{dim = 1}, data = 5,
// legacy num/channels/height/width = 1..4} (caffe/src/caffe/proto/caffe.proto:10-22,64-96,310-380).
#pragma once
#include <map>
#include <string>
#include <vector>

namespace sivo {

struct Blob {
  std::vector<int64_t> shape;
  std::vector<float> data;
  size_t count() const {
    size_t n = 1;
    for (auto d : shape) n *= static_cast<size_t>(d);
    return n;
  }
};

using WeightMap = std::map<std::string, std::vector<Blob>>;  // layer name -> blobs

WeightMap read_caffemodel(const std::string& path);

}  // namespace sivo
