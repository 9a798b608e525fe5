// TEST INFRASTRUCTURE ONLY.  Harness around the reference's *own* expert modules.
//
// Compiled together with /root/reference/core/parallel/expert_module.cpp (the reference source is compiled where it
// lies, never copied) into oracle/_ref/ref_expert_module*.so by oracle/ref_build/Makefile.  It gives the tests the real
// arithmetic of D1/D2/D3: `MixtralMoEDenseActDense::forward` (expert_module.cpp:147-175),
// `DeepSeekMoEDenseActDense::forward` (:193-203), `SwitchTransformersDenseActDense` (:24-35), `...GatedActDense`
// (:54-59), `NllbMoeDenseActDense` (:79-93), `FSGPTMoEDenseActDense` (:113-129), bound to their weights the way the
// reference binds them (`SetTensorsFromBlob` through the global tensor index, :16-22, :45-51, :69-77, :139-145, ...).
//
// The only symbol the reference TU needs from the rest of its engine is the global `kTensorIndex`
// (aio/archer_tensor_index.h:55; defined in prefetch/archer_prefetch_handle.cpp:22) -- defined here.
#include <torch/extension.h>

#include <memory>
#include <vector>

#include "aio/archer_tensor_index.h"
#include "parallel/expert_module.h"

std::unique_ptr<ArcherTensorIndex> kTensorIndex = std::make_unique<ArcherTensorIndex>();

namespace {

template <class M>
torch::Tensor run(int dtype, const std::vector<std::uint32_t>& ids, const torch::Tensor& x) {
  M module(dtype);
  module.SetTensorsFromBlob(nullptr, ids, torch::Device(torch::kCPU));
  return module.forward(x);
}

// expert_type / dtype: the integers of expert_module.h:13-23.  `tensors` in the reference's tensor-id order.
torch::Tensor expert_forward(int expert_type, int dtype, std::vector<torch::Tensor> tensors, torch::Tensor x) {
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

}  // namespace

PYBIND11_MODULE(ref_expert_module, m) {
  m.doc() = "reference expert modules (core/parallel/expert_module.cpp) compiled as-is; test oracle only";
  m.def("expert_forward", &expert_forward, "run the reference's expert module of the given type on CPU");
}
