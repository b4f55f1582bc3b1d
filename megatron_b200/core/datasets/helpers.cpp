// Native index builders for the dataset layer (CPU, pybind11).
//
// Capability parity with the reference's `megatron/core/datasets/helpers.cpp` exports
// (`build_sample_idx_int32/int64`, `build_blending_indices`, `build_exhaustive_blending_indices`,
// `build_mapping`, `build_blocks_mapping`), written from the documented semantics:
//
//  * sample index  : walk the shuffled document order and cut the token stream into samples of
//                    `seq_length + add_extra_token` tokens; consecutive samples overlap by the extra
//                    token.  Row i = (index into document_idx, token offset inside that document) of
//                    the FIRST token of sample i; row i+1 marks the end.
//  * blending      : at step i choose the dataset whose realised share lags its target weight most.
//  * sentence maps : pack consecutive sentences of a document into samples of <= max_seq_length
//                    tokens with a (seeded) probability of producing short samples.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <stdexcept>
#include <vector>

namespace py = pybind11;

template <typename OutT>
py::array build_sample_idx_impl(const py::array_t<int32_t>& sizes_, const py::array_t<int32_t>& document_idx_, const int32_t seq_length, const int32_t num_epochs,
                                const int64_t tokens_per_epoch, const bool drop_last_partial_sequence, const int add_extra_token_to_sequence) {
  auto sizes = sizes_.unchecked<1>();
  auto doc_idx = document_idx_.unchecked<1>();
  if (seq_length <= 1 || num_epochs <= 0 || tokens_per_epoch <= 1) throw std::invalid_argument("build_sample_idx: bad arguments");
  const int64_t total_tokens = (int64_t)num_epochs * tokens_per_epoch - add_extra_token_to_sequence;
  int64_t num_samples = drop_last_partial_sequence ? total_tokens / seq_length : (int64_t)std::ceil((double)total_tokens / seq_length);
  OutT* out = new OutT[2 * (num_samples + 1)];
  int64_t cursor = 0;  // position in document_idx
  OutT offset = 0;     // token offset inside the current document
  out[0] = 0;
  out[1] = 0;
  const int64_t ndocs = doc_idx.shape(0);
  for (int64_t s = 1; s <= num_samples; ++s) {
    int64_t remaining = (int64_t)seq_length + add_extra_token_to_sequence;
    while (remaining != 0) {
      if (cursor >= ndocs) {  // ran out of tokens: only legal for the final, partial sample
        if (drop_last_partial_sequence) {
          delete[] out;
          throw std::runtime_error("build_sample_idx: document index exhausted");
        }
        break;
      }
      const int64_t doc_len = (int64_t)sizes[doc_idx[cursor]] - offset;
      remaining -= doc_len;
      if (remaining <= 0) {
        // sample ends inside this document; the extra token is shared with the next sample
        offset = (OutT)(offset + remaining + doc_len - add_extra_token_to_sequence);
        remaining = 0;
      } else {
        if (cursor == ndocs - 1) {  // last document consumed, keep the boundary at its end
          offset = (OutT)(sizes[doc_idx[cursor]] - add_extra_token_to_sequence);
          break;
        }
        ++cursor;
        offset = 0;
      }
    }
    out[2 * s] = (OutT)cursor;
    out[2 * s + 1] = offset;
  }
  py::capsule owner(out, [](void* p) { delete[] reinterpret_cast<OutT*>(p); });
  return py::array_t<OutT>({num_samples + 1, (int64_t)2}, {2 * sizeof(OutT), sizeof(OutT)}, out, owner);
}

py::array build_sample_idx_int32(const py::array_t<int32_t>& sizes, const py::array_t<int32_t>& document_idx, int32_t seq_length, int32_t num_epochs,
                                 int64_t tokens_per_epoch, bool drop_last_partial_sequence = true, int add_extra_token_to_sequence = 1) {
  return build_sample_idx_impl<int32_t>(sizes, document_idx, seq_length, num_epochs, tokens_per_epoch, drop_last_partial_sequence, add_extra_token_to_sequence);
}
py::array build_sample_idx_int64(const py::array_t<int32_t>& sizes, const py::array_t<int32_t>& document_idx, int32_t seq_length, int32_t num_epochs,
                                 int64_t tokens_per_epoch, bool drop_last_partial_sequence = true, int add_extra_token_to_sequence = 1) {
  return build_sample_idx_impl<int64_t>(sizes, document_idx, seq_length, num_epochs, tokens_per_epoch, drop_last_partial_sequence, add_extra_token_to_sequence);
}

// dataset_index[i] = which dataset sample i comes from, dataset_sample_index[i] = running index inside it
void build_blending_indices(py::array_t<int16_t>& dataset_index, py::array_t<int64_t>& dataset_sample_index, const py::array_t<double>& weights,
                            const int32_t num_datasets, const int64_t size, const bool verbose) {
  auto di = dataset_index.mutable_unchecked<1>();
  auto dsi = dataset_sample_index.mutable_unchecked<1>();
  auto w = weights.unchecked<1>();
  std::vector<int64_t> taken(num_datasets, 0);
  for (int64_t i = 0; i < size; ++i) {
    const double denom = std::max<double>((double)i, 1.0);
    int best = 0;
    double best_err = -std::numeric_limits<double>::infinity();
    for (int d = 0; d < num_datasets; ++d) {
      const double err = w[d] * denom - (double)taken[d];
      if (err > best_err) {
        best_err = err;
        best = d;
      }
    }
    di[i] = (int16_t)best;
    dsi[i] = taken[best]++;
  }
  if (verbose) {
    py::print("> sample ratios (target / achieved):");
    for (int d = 0; d < num_datasets; ++d) py::print("   dataset", d, w[d], (double)taken[d] / (double)std::max<int64_t>(size, 1));
  }
}

// draw from every dataset in proportion to its size until all are exhausted exactly once
void build_exhaustive_blending_indices(py::array_t<int16_t>& dataset_index, py::array_t<int64_t>& dataset_sample_index, const py::array_t<int64_t>& sizes,
                                       const int32_t num_datasets) {
  auto di = dataset_index.mutable_unchecked<1>();
  auto dsi = dataset_sample_index.mutable_unchecked<1>();
  auto sz = sizes.unchecked<1>();
  int64_t total = 0;
  for (int d = 0; d < num_datasets; ++d) total += sz[d];
  std::vector<int64_t> taken(num_datasets, 0);
  for (int64_t i = 0; i < total; ++i) {
    int best = -1;
    double best_err = -std::numeric_limits<double>::infinity();
    for (int d = 0; d < num_datasets; ++d) {
      if (taken[d] >= sz[d]) continue;
      const double err = ((double)sz[d] / (double)total) * (double)std::max<int64_t>(i, 1) - (double)taken[d];
      if (err > best_err) {
        best_err = err;
        best = d;
      }
    }
    di[i] = (int16_t)best;
    dsi[i] = taken[best]++;
  }
}

// BERT/T5: (start sentence, end sentence, target sequence length) triples per sample
static inline int32_t target_len(int32_t short_prob_denominator, int32_t max_length, std::mt19937& gen) {
  if (short_prob_denominator == 0) return max_length;
  const uint32_t r = gen();
  if (r % (uint32_t)short_prob_denominator == 0) return 2 + (int32_t)(r % (uint32_t)(max_length - 1));
  return max_length;
}

py::array build_mapping(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_, const int32_t num_epochs, const uint64_t max_num_samples,
                        const int32_t max_seq_length, const double short_seq_prob, const int32_t seed, const bool verbose, const int32_t min_num_sent) {
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  const int32_t short_den = short_seq_prob > 0 ? (int32_t)std::lround(1.0 / short_seq_prob) : 0;
  std::vector<int64_t> rows;
  for (int pass = 0; pass < 1; ++pass) {
    std::mt19937 gen((uint32_t)seed);
    for (int32_t epoch = 0; epoch < num_epochs && rows.size() / 3 < max_num_samples; ++epoch) {
      for (int64_t d = 0; d + 1 < docs.shape(0) && rows.size() / 3 < max_num_samples; ++d) {
        const int64_t first = docs[d], last = docs[d + 1];
        if (last - first < min_num_sent) continue;
        bool too_long = false;
        for (int64_t s = first; s < last; ++s) too_long |= sizes[s] > 512 * 1024;
        if (too_long) continue;
        int64_t start = first;
        int32_t seq_len = 0, nsent = 0;
        int32_t tgt = target_len(short_den, max_seq_length, gen);
        for (int64_t s = first; s < last; ++s) {
          seq_len += sizes[s];
          ++nsent;
          const int64_t remain = last - s - 1;
          if ((seq_len >= tgt && remain >= min_num_sent && nsent >= min_num_sent) || remain == 0) {
            if (nsent >= min_num_sent) {
              rows.push_back(start);
              rows.push_back(s + 1);
              rows.push_back(tgt);
            }
            start = s + 1;
            seq_len = 0;
            nsent = 0;
            tgt = target_len(short_den, max_seq_length, gen);
          }
        }
      }
    }
  }
  const int64_t n = (int64_t)rows.size() / 3;
  // deterministic shuffle of the samples
  std::vector<int64_t> perm(n);
  for (int64_t i = 0; i < n; ++i) perm[i] = i;
  std::mt19937_64 g64((uint64_t)seed + 1);
  for (int64_t i = n - 1; i > 0; --i) std::swap(perm[i], perm[(int64_t)(g64() % (uint64_t)(i + 1))]);
  int64_t* out = new int64_t[3 * std::max<int64_t>(n, 1)];
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) out[3 * i + k] = rows[3 * perm[i] + k];
  if (verbose) py::print("> build_mapping produced", n, "samples");
  py::capsule owner(out, [](void* p) { delete[] reinterpret_cast<int64_t*>(p); });
  return py::array_t<int64_t>({n, (int64_t)3}, {3 * sizeof(int64_t), sizeof(int64_t)}, out, owner);
}

// ICT/REALM style block mapping: (start sentence, end sentence, document, block id)
py::array build_blocks_mapping(const py::array_t<int64_t>& docs_, const py::array_t<int32_t>& sizes_, const py::array_t<int32_t>& titles_sizes_,
                               const int32_t num_epochs, const uint64_t max_num_samples, const int32_t max_seq_length, const int32_t seed, const bool verbose,
                               const bool use_one_sent_blocks) {
  auto docs = docs_.unchecked<1>();
  auto sizes = sizes_.unchecked<1>();
  auto titles = titles_sizes_.unchecked<1>();
  const int32_t min_sent = use_one_sent_blocks ? 1 : 2;
  std::vector<int64_t> rows;
  int64_t block_id = 0;
  for (int32_t epoch = 0; epoch < num_epochs && rows.size() / 4 < max_num_samples; ++epoch) {
    block_id = 0;
    for (int64_t d = 0; d + 1 < docs.shape(0) && rows.size() / 4 < max_num_samples; ++d) {
      const int64_t first = docs[d], last = docs[d + 1];
      const int32_t budget = max_seq_length - titles[d];
      if (last - first < min_sent) continue;
      int64_t start = first;
      int32_t seq_len = 0, nsent = 0;
      for (int64_t s = first; s < last; ++s) {
        seq_len += sizes[s];
        ++nsent;
        const int64_t remain = last - s - 1;
        if ((seq_len >= budget && remain >= min_sent && nsent >= min_sent) || remain == 0) {
          rows.push_back(start);
          rows.push_back(s + 1);
          rows.push_back(d);
          rows.push_back(block_id++);
          start = s + 1;
          seq_len = 0;
          nsent = 0;
        }
      }
    }
  }
  const int64_t n = (int64_t)rows.size() / 4;
  std::vector<int64_t> perm(n);
  for (int64_t i = 0; i < n; ++i) perm[i] = i;
  std::mt19937_64 g64((uint64_t)seed + 1);
  for (int64_t i = n - 1; i > 0; --i) std::swap(perm[i], perm[(int64_t)(g64() % (uint64_t)(i + 1))]);
  int64_t* out = new int64_t[4 * std::max<int64_t>(n, 1)];
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k) out[4 * i + k] = rows[4 * perm[i] + k];
  if (verbose) py::print("> build_blocks_mapping produced", n, "blocks");
  py::capsule owner(out, [](void* p) { delete[] reinterpret_cast<int64_t*>(p); });
  return py::array_t<int64_t>({n, (int64_t)4}, {4 * sizeof(int64_t), sizeof(int64_t)}, out, owner);
}

PYBIND11_MODULE(helpers_cpp, m) {
  m.def("build_sample_idx_int32", &build_sample_idx_int32, py::arg("sizes"), py::arg("document_idx"), py::arg("seq_length"), py::arg("num_epochs"),
        py::arg("tokens_per_epoch"), py::arg("drop_last_partial_sequence") = true, py::arg("add_extra_token_to_sequence") = 1);
  m.def("build_sample_idx_int64", &build_sample_idx_int64, py::arg("sizes"), py::arg("document_idx"), py::arg("seq_length"), py::arg("num_epochs"),
        py::arg("tokens_per_epoch"), py::arg("drop_last_partial_sequence") = true, py::arg("add_extra_token_to_sequence") = 1);
  m.def("build_blending_indices", &build_blending_indices);
  m.def("build_exhaustive_blending_indices", &build_exhaustive_blending_indices);
  m.def("build_mapping", &build_mapping);
  m.def("build_blocks_mapping", &build_blocks_mapping);
}
