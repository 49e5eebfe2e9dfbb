// GzipTextWriter.h -- ordered, optionally multi-threaded gzip writer for the FASTQ text the device assembles.
//
// threads == 1: the zlib call sequence of the reference's FastqOutfileStream (src/cppNGS/FastqFileStream.cpp:160-172: gzopen "wb",
//   gzbuffer 131072, gzsetparams) fed by one writer thread per file, like the reference's FastqWriter threads
//   (src/SeqPurge/OutputWorker.cpp:24-32) -- the .gz bytes equal the reference's for equal records.
// bgzf: the output is written as BGZF (blocked gzip as htslib/bgzip write it: members of at most 64 KiB that carry their compressed
//   size in a 'BC' extra field, plus the empty end-of-file block), with or without a pool. Any gzip reader reads it; this
//   repository's TextSource and htslib inflate it in parallel.
// threads  > 1: the text is cut into pieces that a shared pool deflates independently (raw deflate, Z_SYNC_FLUSH), written in order
//   as ONE gzip member (header, pieces, empty final block, CRC32/ISIZE trailer; CRCs joined with crc32_combine). The decompressed
//   content is identical, the compressed bytes are not (and a little larger: no matches across piece boundaries).
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace seqpurge
{

// fixed pool of threads running queued closures
class WorkerPool
{
public:
	explicit WorkerPool(int threads);
	~WorkerPool();
	void run(std::function<void()> task);

private:
	std::mutex mu_;
	std::condition_variable cv_;
	std::deque<std::function<void()>> tasks_;
	std::vector<std::thread> threads_;
	bool stop_ = false;
};

class GzipTextWriter
{
public:
	// pool == nullptr: serial zlib stream (reference byte sequence)
	GzipTextWriter(const std::string& filename, int compression_level, WorkerPool* pool, bool bgzf = false);
	~GzipTextWriter();
	GzipTextWriter(const GzipTextWriter&) = delete;
	GzipTextWriter& operator=(const GzipTextWriter&) = delete;

	void write(std::vector<uint8_t>&& text); // called from one thread, in output order; blocks when too much is pending
	void write(const uint8_t* text, size_t n); // the same from a buffer the caller keeps (copied once, piece by piece)
	void close();                            // waits for everything to be on disk; throws what the writer thread met

private:
	struct Piece
	{
		std::vector<uint8_t> text, comp;
		uint32_t crc = 0;
		bool done = false;
	};
	void enqueue(std::vector<uint8_t>&& text);
	void writerLoop();
	void compressPiece(Piece* p);
	void compressPieceBgzf(Piece* p);

	std::string filename_;
	int level_;
	WorkerPool* pool_;
	bool bgzf_ = false;
	std::mutex mu_;
	std::condition_variable cv_;
	std::deque<std::unique_ptr<Piece>> queue_; // in output order
	size_t pending_bytes_ = 0;
	bool closing_ = false, closed_ = false;
	std::exception_ptr failure_;
	std::thread writer_;
};

} // namespace seqpurge
