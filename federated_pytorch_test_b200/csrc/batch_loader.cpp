// Native host-side batch assembler (runtime component, no CUDA dependency).
//
// Role: the reference feeds its training loop through a torch DataLoader with one
// worker process that decodes PIL images and normalises them on the CPU, then a
// synchronous .to(device) (/root/reference/src/federated_multi.py:83-85,175).  Here the
// dataset is a dense uint8 array in (pinned) host memory; a small pool of native
// threads gathers the rows of the NEXT batches into a ring of pinned staging slots
// while the GPU computes, so the training thread only ever issues one async H2D copy
// per batch.  Normalisation happens on the device (fused kernel), not here.
//
// Threading model: one coordinator thread walks the epoch's index order and fills
// free slots in order; each fill is split across `n_threads` helpers by row range.
// The consumer calls acquire() (blocks until the next slot in sequence is full) and
// release(slot) when the H2D copy that reads the slot has been enqueued+recorded.
#include <torch/extension.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Slot {
  uint8_t* images = nullptr;   // [batch, row_bytes] pinned
  int64_t* labels = nullptr;   // [batch] pinned
  int64_t count = 0;           // rows valid
  int64_t batch_id = -1;       // which batch of the epoch is stored here
  bool full = false;
};

class BatchAssembler {
 public:
  BatchAssembler(torch::Tensor images, torch::Tensor labels, int64_t batch_size,
                 std::vector<torch::Tensor> slot_images, std::vector<torch::Tensor> slot_labels,
                 int64_t n_threads)
      : images_(images.contiguous()), labels_(labels.contiguous()), batch_(batch_size),
        n_threads_(std::max<int64_t>(1, n_threads)) {
    TORCH_CHECK(!images_.is_cuda() && !labels_.is_cuda(), "dataset must live in host memory");
    TORCH_CHECK(images_.scalar_type() == torch::kUInt8, "images must be uint8");
    TORCH_CHECK(labels_.scalar_type() == torch::kInt64, "labels must be int64");
    TORCH_CHECK(slot_images.size() == slot_labels.size() && !slot_images.empty(), "need >=1 staging slot");
    n_rows_ = images_.size(0);
    row_bytes_ = images_.numel() / std::max<int64_t>(1, n_rows_);
    keep_ = slot_images;  // keep staging tensors alive
    keep_.insert(keep_.end(), slot_labels.begin(), slot_labels.end());
    for (size_t i = 0; i < slot_images.size(); ++i) {
      TORCH_CHECK(slot_images[i].numel() >= batch_ * row_bytes_, "staging slot too small");
      Slot s;
      s.images = slot_images[i].data_ptr<uint8_t>();
      s.labels = slot_labels[i].data_ptr<int64_t>();
      slots_.push_back(s);
    }
    worker_ = std::thread([this] { this->run(); });
  }

  ~BatchAssembler() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (worker_.joinable()) worker_.join();
  }

  // Begin a new epoch over `order` (int64 row indices).  Any unconsumed batches are dropped.
  void start_epoch(torch::Tensor order) {
    auto o = order.to(torch::kInt64).contiguous();
    std::lock_guard<std::mutex> lk(mu_);
    order_ = o;
    n_batches_ = (o.numel() + batch_ - 1) / batch_;
    next_fill_ = 0;
    next_take_ = 0;
    ++epoch_;
    for (auto& s : slots_) { s.full = false; s.batch_id = -1; }
    cv_.notify_all();
  }

  // Blocks until the next batch in sequence is staged.  Returns {slot, rows}; slot = -1 at epoch end.
  std::pair<int64_t, int64_t> acquire() {
    py::gil_scoped_release nogil;
    std::unique_lock<std::mutex> lk(mu_);
    if (next_take_ >= n_batches_) return {-1, 0};
    const int64_t want = next_take_;
    const int64_t slot = want % static_cast<int64_t>(slots_.size());
    cv_.wait(lk, [&] { return stop_ || (slots_[slot].full && slots_[slot].batch_id == want); });
    if (stop_) return {-1, 0};
    ++next_take_;
    return {slot, slots_[slot].count};
  }

  void release(int64_t slot) {
    std::lock_guard<std::mutex> lk(mu_);
    slots_[slot].full = false;
    slots_[slot].batch_id = -1;
    cv_.notify_all();
  }

  int64_t num_batches() const { return n_batches_; }
  int64_t rows_assembled() const { return rows_done_.load(); }

 private:
  void gather(Slot& s, const int64_t* idx, int64_t count) {
    const uint8_t* src = images_.data_ptr<uint8_t>();
    const int64_t* lab = labels_.data_ptr<int64_t>();
    auto body = [&](int64_t a, int64_t b) {
      for (int64_t r = a; r < b; ++r) {
        const int64_t j = idx[r];
        std::memcpy(s.images + r * row_bytes_, src + j * row_bytes_, static_cast<size_t>(row_bytes_));
        s.labels[r] = lab[j];
      }
    };
    if (n_threads_ == 1 || count < 64) {
      body(0, count);
    } else {
      std::vector<std::thread> pool;
      const int64_t per = (count + n_threads_ - 1) / n_threads_;
      for (int64_t t = 1; t < n_threads_; ++t) {
        const int64_t a = t * per, b = std::min(count, a + per);
        if (a < b) pool.emplace_back(body, a, b);
      }
      body(0, std::min(count, per));
      for (auto& th : pool) th.join();
    }
    rows_done_ += count;
  }

  void run() {
    std::unique_lock<std::mutex> lk(mu_);
    while (!stop_) {
      const int64_t nslots = static_cast<int64_t>(slots_.size());
      const int64_t b = next_fill_;
      const bool can = (b < n_batches_) && !slots_[b % nslots].full && (b - next_take_ < nslots);
      if (!can) {
        cv_.wait(lk);
        continue;
      }
      Slot& s = slots_[b % nslots];
      const int64_t my_epoch = epoch_;
      torch::Tensor order = order_;  // keep alive outside the lock
      const int64_t a = b * batch_;
      const int64_t count = std::min<int64_t>(batch_, order.numel() - a);
      lk.unlock();
      gather(s, order.data_ptr<int64_t>() + a, count);
      lk.lock();
      if (my_epoch != epoch_) continue;  // epoch restarted while we were copying
      s.count = count;
      s.batch_id = b;
      s.full = true;
      ++next_fill_;
      cv_.notify_all();
    }
  }

  torch::Tensor images_, labels_, order_;
  std::vector<torch::Tensor> keep_;
  std::vector<Slot> slots_;
  int64_t batch_, n_threads_, n_rows_ = 0, row_bytes_ = 0;
  int64_t n_batches_ = 0, next_fill_ = 0, next_take_ = 0, epoch_ = 0;
  std::atomic<int64_t> rows_done_{0};
  bool stop_ = false;
  std::mutex mu_;
  std::condition_variable cv_;
  std::thread worker_;
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "native batch assembler for pinned-host datasets";
  py::class_<BatchAssembler>(m, "BatchAssembler")
      .def(py::init<torch::Tensor, torch::Tensor, int64_t, std::vector<torch::Tensor>, std::vector<torch::Tensor>, int64_t>())
      .def("start_epoch", &BatchAssembler::start_epoch)
      .def("acquire", &BatchAssembler::acquire)
      .def("release", &BatchAssembler::release)
      .def("num_batches", &BatchAssembler::num_batches)
      .def("rows_assembled", &BatchAssembler::rows_assembled);
}
