// TEST INFRASTRUCTURE ONLY.  Harness around the reference's *own* expert modules.
//
// Compiled together with /root/reference/core/parallel/expert_module.cpp (the reference source is compiled where it
// lies, never copied) into oracle/_ref/ref_expert_module*.so by oracle/ref_build/Makefile.  It gives the tests the real
// arithmetic of D1/D2/D3: `MixtralMoEDenseActDense::forward` (expert_module.cpp:147-175),
// `DeepSeekMoEDenseActDense::forward` (:193-203), `SwitchTransformersDenseActDense` (:24-35), `...GatedActDense`
// (:54-59), `NllbMoeDenseActDense` (:79-93), `FSGPTMoEDenseActDense` (:113-129), bound to their weights the way the
// reference binds them (`SetTensorsFromBlob` through the global tensor index, :16-22, :45-51, :69-77, :139-145, ...).
//
// The only symbol that TU needs from the rest of the engine is the global `kTensorIndex` (aio/archer_tensor_index.h:55),
// which the reference defines in aio/archer_tensor_index.cpp:103 -- that file is compiled in as well (as-is) and also
// provides the reference's on-disk index format (`ArcherTensorIndex::Serialize/Deserialize`, :105-132), exposed below
// for tests/test_store_format.py.
#include <torch/extension.h>

#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "aio/archer_tensor_index.h"
#include "parallel/expert_module.h"


namespace {

template <class M>
torch::Tensor run(int dtype, const std::vector<std::uint32_t>& ids, const torch::Tensor& x) {
  M module(dtype);
  module.SetTensorsFromBlob(nullptr, ids, torch::Device(torch::kCPU));
  return module.forward(x);
}

// expert_type / dtype: the integers of expert_module.h:13-23.  `tensors` in the reference's tensor-id order.
torch::Tensor expert_forward(int expert_type, int dtype, std::vector<torch::Tensor> tensors, torch::Tensor x) {
  if (!kTensorIndex) kTensorIndex = std::make_unique<ArcherTensorIndex>();   // archer_prefetch_handle.cpp:22
  kTensorIndex->clear();
  std::vector<std::uint32_t> ids;
  for (size_t i = 0; i < tensors.size(); ++i) {
    TensorStorageMeta meta;
    meta.id = static_cast<TensorID>(i);
    meta.tensor = tensors[i];
    kTensorIndex->emplace(static_cast<std::uint32_t>(i), meta);
    ids.push_back(static_cast<std::uint32_t>(i));
  }
  torch::NoGradGuard no_grad;
  switch (expert_type) {
    case SWITCH_TRANSFORMERS_DENSE_ACT_DENSE: return run<SwitchTransformersDenseActDense>(dtype, ids, x);
    case SWITCH_TRANSFORMERS_DENSE_GATED_ACT_DENSE: return run<SwitchTransformersDenseGatedActDense>(dtype, ids, x);
    case NLLB_MOE_DENSE_ACT_DENSE: return run<NllbMoeDenseActDense>(dtype, ids, x);
    case FSGPT_MOE_DENSE_ACT_DENSE: return run<FSGPTMoEDenseActDense>(dtype, ids, x);
    case MIXTRAL_MOE_DENSE_ACT_DENSE: return run<MixtralMoEDenseActDense>(dtype, ids, x);
    case DEEPSEEK_MOE_DENSE_ACT_DENSE: return run<DeepSeekMoEDenseActDense>(dtype, ids, x);
    default: throw std::invalid_argument("unknown expert_type");
  }
}

// ---- on-disk index format (aio/archer_tensor_index.cpp:51-132), driven exactly like ArcherTensorHandle::StoreTensor
// fills it (archer_tensor_handle.cpp:53-86): meta = {file_id, offset, nbytes, sizes, options of the stored tensor}
void index_serialize(const std::string& path,
                     const std::vector<std::tuple<std::uint32_t, std::uint32_t, std::int64_t, torch::Tensor>>& entries) {
  ArcherTensorIndex index;
  for (const auto& [id, file_id, offset, t] : entries) {
    TensorStorageMeta meta{file_id, offset, t.nbytes(), t.sizes().vec()};
    meta.options = t.options();
    meta.id = id;
    index.insert(std::make_pair(id, meta));
  }
  index.Serialize(path.c_str());
}

// -> [(id, file_id, offset, size, shape, scalar_type, device_type, device_index, layout, requires_grad, pinned)]
std::vector<std::tuple<std::uint32_t, std::uint32_t, std::int64_t, std::uint64_t, std::vector<std::int64_t>, int, int, int, int,
                       bool, bool>>
index_deserialize(const std::string& path) {
  ArcherTensorIndex index;
  index.Deserialize(path.c_str());
  std::vector<std::tuple<std::uint32_t, std::uint32_t, std::int64_t, std::uint64_t, std::vector<std::int64_t>, int, int, int,
                         int, bool, bool>> out;
  for (const auto& [id, m] : index)
    out.emplace_back(id, m.file_id, m.offset, (std::uint64_t)m.size, m.shape, (int)m.options.dtype().toScalarType(),
                     (int)m.options.device().type(), (int)m.options.device().index(), (int)m.options.layout(),
                     m.options.requires_grad(), m.options.pinned_memory());
  return out;
}

}  // namespace

PYBIND11_MODULE(ref_expert_module, m) {
  m.doc() = "reference expert modules (core/parallel/expert_module.cpp) compiled as-is; test oracle only";
  m.def("expert_forward", &expert_forward, "run the reference's expert module of the given type on CPU");
  m.def("index_serialize", &index_serialize, "ArcherTensorIndex::Serialize of metas built like StoreTensor builds them");
  m.def("index_deserialize", &index_deserialize, "ArcherTensorIndex::Deserialize -> list of meta tuples");
}
