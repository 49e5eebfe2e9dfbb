// ChunkReader.h -- the reader side of the text pipelines (seqpurge_b200's stream pipeline, readqc_b200): one thread per input file list
// inflates (TextSource) and cuts the text after every `pairs` records, the way InputWorker::run fills a job
// (src/SeqPurge/InputWorker.cpp:16-77); the chunks travel through a bounded queue.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <exception>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "SeqPurgeTypes.h"

namespace seqpurge
{

class WorkerPool;

struct TextChunk
{
	std::vector<uint8_t> data;
	int records = 0;      // entries readEntry would deliver for this text (an unterminated last line and an incomplete last record count)
	int max_read_len = 0; // longest bases/qualities line (may include trailing '\r')
	bool file_end = false;
	size_t file_index = 0;
};

class ChunkQueue
{
public:
	explicit ChunkQueue(size_t depth) : depth_(depth) {}
	void push(std::unique_ptr<TextChunk> c)
	{
		std::unique_lock<std::mutex> l(mu_);
		cv_.wait(l, [this] { return q_.size() < depth_ || aborted_; });
		if (aborted_) return;
		q_.push_back(std::move(c));
		cv_.notify_all();
	}
	std::unique_ptr<TextChunk> pop() // nullptr: the reader is done (or failed: see failure())
	{
		std::unique_lock<std::mutex> l(mu_);
		cv_.wait(l, [this] { return !q_.empty() || done_ || aborted_; });
		if (q_.empty()) return nullptr;
		std::unique_ptr<TextChunk> c = std::move(q_.front());
		q_.pop_front();
		cv_.notify_all();
		return c;
	}
	void finish(std::exception_ptr e)
	{
		std::lock_guard<std::mutex> g(mu_);
		done_ = true;
		failure_ = e;
		cv_.notify_all();
	}
	void abort()
	{
		std::lock_guard<std::mutex> g(mu_);
		aborted_ = true;
		cv_.notify_all();
	}
	std::exception_ptr failure()
	{
		std::lock_guard<std::mutex> g(mu_);
		return failure_;
	}

private:
	size_t depth_;
	std::mutex mu_;
	std::condition_variable cv_;
	std::deque<std::unique_ptr<TextChunk>> q_;
	bool done_ = false, aborted_ = false;
	std::exception_ptr failure_;
};

// reads the files of one list, cuts the inflated text after every `pairs` records; BGZF files are inflated by the pool (may be null)
void readerLoop(const std::vector<std::string>& files, int pairs, ChunkQueue& out, WorkerPool* pool);

// the `index`-th record of a chunk as the reference's reader delivers it (error reporting only)
FastqEntry entryAt(const TextChunk& c, int index);

} // namespace seqpurge
