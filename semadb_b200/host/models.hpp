// models.hpp — the parameter / option / result structs the hot path honours, with the
// reference's validation rules and messages (models/index.go:275-313, models/quantizer.go:5-76,
// models/search.go:238-306, models/constants.go).
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "diskstore.hpp"

namespace semadb {
namespace models {

constexpr const char* DistanceEuclidean = "euclidean";
constexpr const char* DistanceCosine = "cosine";
constexpr const char* DistanceDot = "dot";
constexpr const char* DistanceHamming = "hamming";
constexpr const char* DistanceJaccard = "jaccard";
constexpr const char* DistanceHaversine = "haversine";
constexpr const char* QuantizerNone = "none";
constexpr const char* QuantizerBinary = "binary";
constexpr const char* QuantizerProduct = "product";
constexpr const char* OperatorNear = "near";

struct BinaryQuantizerParamaters {  // models/quantizer.go:30-39 (spelling as in the reference)
  std::optional<float> Threshold;
  int TriggerThreshold = 0;
  std::string DistanceMetric = DistanceHamming;
  Error Validate() const {  // quantizer.go:41-49
    if (!Threshold && (TriggerThreshold < 0 || TriggerThreshold > 50000))
      return Error("triggerThreshold must be between 0 and 50000, got " + std::to_string(TriggerThreshold));
    if (DistanceMetric != DistanceHamming && DistanceMetric != DistanceJaccard)
      return Error("invalid distance metric for binary quantization, got " + DistanceMetric);
    return Ok();
  }
};

struct ProductQuantizerParameters {  // models/quantizer.go:52-63
  int NumCentroids = 256;
  int NumSubVectors = 8;
  int TriggerThreshold = 10000;
  Error Validate() const {  // quantizer.go:65-76
    if (NumCentroids < 2 || NumCentroids > 256)
      return Error("numCentroids must be between 2 and 256, got " + std::to_string(NumCentroids));
    if (NumSubVectors < 2) return Error("numSubVectors must be at least 2, got " + std::to_string(NumSubVectors));
    if (TriggerThreshold < 1000 || TriggerThreshold > 10000)
      return Error("triggerThreshold must be between 1000 and 10000, got " + std::to_string(TriggerThreshold));
    return Ok();
  }
};

struct Quantizer {  // models/quantizer.go:5-28
  std::string Type = QuantizerNone;
  std::optional<BinaryQuantizerParamaters> Binary;
  std::optional<ProductQuantizerParameters> Product;
  Error Validate() const {
    if (Type == QuantizerNone) return Ok();
    if (Type == QuantizerBinary) {
      if (!Binary) return Error("binary quantizer parameters not provided");
      return Binary->Validate();
    }
    if (Type == QuantizerProduct) {
      if (!Product) return Error("product quantizer parameters not provided");
      return Product->Validate();
    }
    return Error("unknown quantizer type " + Type);
  }
};

struct IndexVectorVamanaParameters {  // models/index.go:275-282
  unsigned VectorSize = 0;
  std::string DistanceMetric = DistanceEuclidean;
  int SearchSize = 75;
  int DegreeBound = 64;
  float Alpha = 1.2f;
  std::optional<Quantizer> Quantizer_;
  Error Validate() const {  // index.go:284-313
    if (VectorSize < 1 || VectorSize > 4096)
      return Error("vector size must be between 1 and 4096, got " + std::to_string(VectorSize));
    const std::string& m = DistanceMetric;
    if (m != DistanceEuclidean && m != DistanceCosine && m != DistanceDot && m != DistanceHamming && m != DistanceJaccard &&
        m != DistanceHaversine)
      return Error("unknown distance metric " + m);
    if (m == DistanceHaversine && VectorSize != 2)
      return Error("haversine distance metric requires vector size 2 got " + std::to_string(VectorSize));
    if (SearchSize < 25 || SearchSize > 75)
      return Error("search size must be between 25 and 75, got " + std::to_string(SearchSize));
    if (DegreeBound < 32 || DegreeBound > 64)
      return Error("degree bound must be between 32 and 64, got " + std::to_string(DegreeBound));
    if (Alpha < 1.1f || Alpha > 1.5f) return Error("alpha must be between 1.1 and 1.5, got " + std::to_string(Alpha));
    if (Quantizer_) return Quantizer_->Validate();
    return Ok();
  }
};

struct SearchVectorVamanaOptions {  // models/search.go:268-275
  std::vector<float> Vector;
  std::string Operator = OperatorNear;
  int SearchSize = 75;
  int Limit = 10;
  std::optional<float> Weight;
  Error Validate() const {  // search.go:277-306
    if (Vector.size() < 1 || Vector.size() > 4096)
      return Error("query vector length must be between 1 and 4096, got " + std::to_string(Vector.size()));
    if (Operator != OperatorNear)
      return Error("invalid operator " + Operator + " for vector query, expected " + OperatorNear);
    if (SearchSize < 25 || SearchSize > 75)
      return Error("invalid searchSize " + std::to_string(SearchSize) + " for vector query, expected 25-75");
    if (Limit < 1 || Limit > 75) return Error("invalid limit " + std::to_string(Limit) + " for vector query, expected 1-75");
    if (SearchSize < Limit) return Error("searchSize must be greater than or equal to limit");
    return Ok();
  }
};

struct SearchResult {  // models/search.go:238-252 (the fields the vector index fills)
  uint64_t NodeId = 0;
  float Distance = 0;
  float HybridScore = 0;
};

}  // namespace models
}  // namespace semadb
