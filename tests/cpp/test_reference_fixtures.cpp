// C++ host-mirror test: the reference's own integration fixtures (tests/test.rs of SeekStorm) through ssb::Index::search.
//   lexical: 4 docs, field `body` = "body1", "body1", "body2 test", "body3 test"            (tests/test.rs:64-86)
//            AND "+body2 +test" -> 1 result / count 1 / total 1; Union Count "test" -> 0 results, total 2   (:150-208)
//   vector : 3 x 128-d f32 Euclidean, v_j[i] = 0.001*(128 j + i + 1), AnnMode::All, length 10 -> 3 results     (:675-745)
// exit codes: 0 = all checks passed, 3 = no CUDA device (expected on the CPU-only build box), 1 = failure.
#include <cmath>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "seekstorm_b200.hpp"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); fails++; } } while (0)

// index.rs:4237-4251 int_to_byte4 for tiny lengths (< 24: identity)
static uint8_t byte4(uint32_t len) { return (uint8_t)len; }

int main() {
    try {
        // ---------------- lexical fixture ----------------
        ssb::Index ix(0, 128, ssb::VectorSimilarity::Euclidean);
        const std::vector<std::vector<std::string>> docs = {{"body1"}, {"body1"}, {"body2", "test"}, {"body3", "test"}};
        std::map<std::string, std::vector<std::pair<uint16_t, uint16_t>>> post;   // term -> (doc, tf), docs ascending
        std::vector<uint8_t> len_bytes; uint64_t len_sum = 0;
        for (size_t d = 0; d < docs.size(); d++) {
            for (auto& t : docs[d]) {
                auto& p = post[t];
                if (!p.empty() && p.back().first == d) p.back().second++; else p.push_back({(uint16_t)d, 1});
            }
            len_bytes.push_back(byte4((uint32_t)docs[d].size())); len_sum += docs[d].size();
        }
        std::vector<uint64_t> keys; std::vector<uint32_t> offs{0}; std::vector<uint16_t> ids, tfs;
        for (auto& kv : post) {
            keys.push_back(ssb::fnv1a64(kv.first));
            for (auto& p : kv.second) { ids.push_back(p.first); tfs.push_back(p.second); }
            offs.push_back((uint32_t)ids.size());
        }
        ssb_level_desc lv{0, 4, (uint32_t)keys.size(), 0, keys.data(), offs.data(), ids.data(), tfs.data(), len_bytes.data()};
        ix.add_lexical_level(lv);
        ix.commit(4, len_sum);
        auto ro = ix.search("+body2 +test", std::nullopt, ssb::QueryType::Intersection, ssb::SearchMode::lexical(), false, 0, 10, ssb::ResultType::TopkCount);
        CHECK(ro.results.size() == 1); CHECK(ro.result_count == 1); CHECK(ro.result_count_total == 1);
        CHECK(!ro.results.empty() && ro.results[0].doc_id == 2);
        // hand-computed (tests/golden/golden.json and_body2_test): idf(4,1)*2.2/(1+cache) + idf(4,2)*2.2/(1+cache)
        CHECK(!ro.results.empty() && std::fabs(ro.results[0].score - 1.6694655418395996f) < 1e-6f);
        ro = ix.search("test", std::nullopt, ssb::QueryType::Union, ssb::SearchMode::lexical(), false, 0, 10, ssb::ResultType::Count);
        CHECK(ro.results.empty()); CHECK(ro.result_count == 0); CHECK(ro.result_count_total == 2);
        ro = ix.search("body2 test", std::nullopt, ssb::QueryType::Union, ssb::SearchMode::lexical(), false, 0, 10, ssb::ResultType::TopkCount);
        CHECK(ro.results.size() == 2 && ro.results[0].doc_id == 2 && ro.results[1].doc_id == 3 && ro.result_count_total == 2);
        ro = ix.search("nosuchterm", std::nullopt, ssb::QueryType::Union, ssb::SearchMode::lexical(), false, 0, 10, ssb::ResultType::TopkCount);
        CHECK(ro.results.empty() && ro.result_count_total == 0);           // infallible: empty ResultObject
        // ---------------- vector fixture ----------------
        std::vector<float> rows(3 * 128);
        for (int j = 0; j < 3; j++) for (int i = 0; i < 128; i++) rows[j * 128 + i] = (float)((128 * j + i + 1) / 1000.0);
        ix.add_vector_level(0, rows.data(), 3, 128);
        std::vector<float> q(rows.begin(), rows.begin() + 128);
        ro = ix.search("", q, ssb::QueryType::Union, ssb::SearchMode::vector(), false, 0, 10, ssb::ResultType::TopkCount);
        CHECK(ro.results.size() == 3); CHECK(ro.result_count == 3); CHECK(ro.result_count_total == 3);
        CHECK(ro.results.size() == 3 && ro.results[0].doc_id == 0 && ro.results[1].doc_id == 1 && ro.results[2].doc_id == 2);
        CHECK(ro.results.size() == 3 && ro.results[0].score == 0.0f && std::fabs(ro.results[1].score - (-2.097152f)) < 1e-4f);
        // unsupported arguments fail loudly instead of being ignored
        bool threw = false;
        try { ix.search("test", std::nullopt, ssb::QueryType::Union, ssb::SearchMode::lexical(), false, 0, 10, ssb::ResultType::Topk, true); }
        catch (const ssb::Error&) { threw = true; }
        CHECK(threw);
    } catch (const ssb::Error& e) {
        if (e.code == SSB_E_NO_DEVICE) { std::printf("no CUDA device: %s\n", e.what()); return 3; }
        std::printf("ssb::Error %d: %s\n", e.code, e.what());
        return 1;
    }
    std::printf(fails ? "FAILED (%d)\n" : "OK\n", fails);
    return fails ? 1 : 0;
}
