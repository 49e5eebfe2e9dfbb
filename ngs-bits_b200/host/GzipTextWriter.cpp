#include "GzipTextWriter.h"

#include <zlib.h>

#include <cstring>

#include "SeqPurgeTypes.h"

namespace seqpurge
{

WorkerPool::WorkerPool(int threads)
{
	for (int i = 0; i < threads; ++i)
	{
		threads_.emplace_back([this]() {
			for (;;)
			{
				std::function<void()> task;
				{
					std::unique_lock<std::mutex> l(mu_);
					cv_.wait(l, [this] { return stop_ || !tasks_.empty(); });
					if (tasks_.empty()) return;
					task = std::move(tasks_.front());
					tasks_.pop_front();
				}
				task();
			}
		});
	}
}

WorkerPool::~WorkerPool()
{
	{
		std::lock_guard<std::mutex> g(mu_);
		stop_ = true;
	}
	cv_.notify_all();
	for (std::thread& t : threads_) t.join();
}

void WorkerPool::run(std::function<void()> task)
{
	{
		std::lock_guard<std::mutex> g(mu_);
		tasks_.push_back(std::move(task));
	}
	cv_.notify_one();
}

namespace
{
constexpr size_t kPieceBytes = 1 << 20;      // text per independently deflated piece
constexpr size_t kMaxPendingBytes = 256u << 20; // back-pressure on the producer
} // namespace

GzipTextWriter::GzipTextWriter(const std::string& filename, int compression_level, WorkerPool* pool, bool bgzf)
    : filename_(filename), level_(compression_level), pool_(pool), bgzf_(bgzf)
{
	if (compression_level < 0 || compression_level > 9)
		throw ArgumentException("Invalid gzip compression level '" + std::to_string(compression_level) + "' given for FASTQ file '" + filename + "'!");
	writer_ = std::thread([this]() { writerLoop(); });
}

GzipTextWriter::~GzipTextWriter()
{
	try
	{
		close();
	}
	catch (...)
	{
	}
}

void GzipTextWriter::write(std::vector<uint8_t>&& text)
{
	if (text.empty()) return;
	if (!pool_ && !bgzf_) enqueue(std::move(text)); // one piece, no copy
	else write(text.data(), text.size());
}

void GzipTextWriter::write(const uint8_t* text, size_t total)
{
	size_t off = 0;
	while (off < total)
	{
		const size_t n = (pool_ || bgzf_) ? std::min(kPieceBytes, total - off) : total;
		enqueue(std::vector<uint8_t>(text + off, text + off + n));
		off += n;
	}
}

void GzipTextWriter::enqueue(std::vector<uint8_t>&& text)
{
	std::unique_ptr<Piece> p(new Piece());
	p->text = std::move(text);
	Piece* raw = p.get();
	{
		std::unique_lock<std::mutex> l(mu_);
		cv_.wait(l, [this] { return pending_bytes_ < kMaxPendingBytes || failure_; });
		if (failure_) std::rethrow_exception(failure_);
		pending_bytes_ += raw->text.size();
		if (!pool_) raw->done = true; // the writer thread compresses in order itself
		queue_.push_back(std::move(p));
	}
	if (pool_) pool_->run([this, raw]() { bgzf_ ? compressPieceBgzf(raw) : compressPiece(raw); });
	else cv_.notify_all();
}

void GzipTextWriter::compressPiece(Piece* p)
{
	try
	{
		z_stream zs;
		memset(&zs, 0, sizeof(zs));
		if (deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw Exception("deflateInit2 failed");
		p->comp.resize(deflateBound(&zs, (uLong)p->text.size()) + 16);
		zs.next_in = p->text.data();
		zs.avail_in = (uInt)p->text.size();
		zs.next_out = p->comp.data();
		zs.avail_out = (uInt)p->comp.size();
		const int rc = deflate(&zs, Z_SYNC_FLUSH); // ends on a byte boundary, no final block
		if (rc != Z_OK || zs.avail_in != 0)
		{
			deflateEnd(&zs);
			throw Exception("deflate failed for '" + filename_ + "'");
		}
		p->comp.resize(p->comp.size() - zs.avail_out);
		deflateEnd(&zs);
		p->crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), p->text.data(), (uInt)p->text.size());
		std::lock_guard<std::mutex> g(mu_);
		p->done = true;
		cv_.notify_all(); // under the lock: once `done` is visible the writer may be closed and destroyed, this task must not touch it afterwards
	}
	catch (...)
	{
		std::lock_guard<std::mutex> g(mu_);
		if (!failure_) failure_ = std::current_exception();
		p->done = true;
		cv_.notify_all();
	}
}

// the piece as a sequence of BGZF blocks (SAM specification, section 4.1): 18 bytes of header with the block size, raw deflate of at
// most 0xff00 bytes of text, CRC32, ISIZE
void GzipTextWriter::compressPieceBgzf(Piece* p)
{
	try
	{
		constexpr size_t kBlockText = 0xff00;
		z_stream zs;
		memset(&zs, 0, sizeof(zs));
		if (deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw Exception("deflateInit2 failed");
		const size_t n_blocks = (p->text.size() + kBlockText - 1) / kBlockText;
		p->comp.resize(n_blocks * 65536);
		size_t out = 0;
		for (size_t off = 0; off < p->text.size(); off += kBlockText)
		{
			const size_t n = std::min(kBlockText, p->text.size() - off);
			uint8_t* b = p->comp.data() + out;
			const uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
			memcpy(b, hdr, 18);
			deflateReset(&zs);
			zs.next_in = p->text.data() + off;
			zs.avail_in = (uInt)n;
			zs.next_out = b + 18;
			zs.avail_out = 65536 - 18 - 8;
			if (deflate(&zs, Z_FINISH) != Z_STREAM_END)
			{
				deflateEnd(&zs);
				throw Exception("deflate failed for '" + filename_ + "'");
			}
			const size_t clen = (65536 - 18 - 8) - zs.avail_out;
			const size_t bsize = 18 + clen + 8;
			b[16] = (uint8_t)((bsize - 1) & 0xff);
			b[17] = (uint8_t)((bsize - 1) >> 8);
			const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), p->text.data() + off, (uInt)n);
			for (int i = 0; i < 4; ++i) b[18 + clen + i] = (uint8_t)(crc >> (8 * i));
			for (int i = 0; i < 4; ++i) b[18 + clen + 4 + i] = (uint8_t)((uint32_t)n >> (8 * i));
			out += bsize;
		}
		deflateEnd(&zs);
		p->comp.resize(out);
	}
	catch (...)
	{
		std::lock_guard<std::mutex> g(mu_);
		if (!failure_) failure_ = std::current_exception();
	}
	{
		std::lock_guard<std::mutex> g(mu_);
		p->done = true;
		cv_.notify_all(); // under the lock, see compressPiece
	}
}

void GzipTextWriter::writerLoop()
{
	gzFile gz = nullptr;
	FILE* fp = nullptr;
	uint32_t crc = (uint32_t)crc32(0L, Z_NULL, 0);
	uint64_t isize = 0;
	try
	{
		if (!pool_ && !bgzf_)
		{
			gz = gzopen(filename_.c_str(), "wb");
			if (!gz) throw FileAccessException("Could not open file '" + filename_ + "' for writing!");
			gzbuffer(gz, 131072);
			gzsetparams(gz, level_, Z_DEFAULT_STRATEGY);
		}
		else
		{
			fp = fopen(filename_.c_str(), "wb");
			if (!fp) throw FileAccessException("Could not open file '" + filename_ + "' for writing!");
			// the header zlib's gzopen("wb") writes: no name, no time, XFL as deflate sets it, OS = Unix
			const unsigned char xfl = level_ == 9 ? 2 : (level_ < 2 ? 4 : 0);
			const unsigned char hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, xfl, 3};
			if (!bgzf_ && fwrite(hdr, 1, 10, fp) != 10) throw FileAccessException("Could not write to file '" + filename_ + "'!");
		}
		for (;;)
		{
			std::unique_ptr<Piece> p;
			{
				std::unique_lock<std::mutex> l(mu_);
				cv_.wait(l, [this] { return (!queue_.empty() && queue_.front()->done) || (queue_.empty() && closing_) || failure_; });
				if (failure_) break;
				if (queue_.empty()) break; // closing
				p = std::move(queue_.front());
				queue_.pop_front();
			}
			if (bgzf_ && !pool_) // serial BGZF: deflate here, in order
			{
				compressPieceBgzf(p.get());
				std::unique_lock<std::mutex> l(mu_);
				if (failure_) break; // the piece's buffer is not valid output: nothing of it is written
			}
			if (gz)
			{
				size_t off = 0; // gzwrite takes an unsigned length
				while (off < p->text.size())
				{
					const unsigned n = (unsigned)std::min<size_t>(p->text.size() - off, 1u << 30);
					if (gzwrite(gz, p->text.data() + off, n) != (int)n) throw FileAccessException("Could not write to file '" + filename_ + "'!");
					off += n;
				}
			}
			else
			{
				if (fwrite(p->comp.data(), 1, p->comp.size(), fp) != p->comp.size()) throw FileAccessException("Could not write to file '" + filename_ + "'!");
				crc = (uint32_t)crc32_combine(crc, p->crc, (z_off_t)p->text.size());
				isize += p->text.size();
			}
			{
				std::lock_guard<std::mutex> g(mu_);
				pending_bytes_ -= p->text.size();
			}
			cv_.notify_all();
		}
		if (gz)
		{
			if (gzclose(gz) != Z_OK) throw FileAccessException("Could not write to file '" + filename_ + "'!");
			gz = nullptr;
		}
		if (fp)
		{
			unsigned char tail[10] = {0x03, 0x00}; // empty static block with BFINAL, then CRC32 and ISIZE (little endian)
			for (int i = 0; i < 4; ++i) tail[2 + i] = (unsigned char)(crc >> (8 * i));
			for (int i = 0; i < 4; ++i) tail[6 + i] = (unsigned char)((uint32_t)isize >> (8 * i));
			static const unsigned char bgzf_eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
			const bool ok = bgzf_ ? fwrite(bgzf_eof, 1, 28, fp) == 28 : fwrite(tail, 1, 10, fp) == 10;
			const bool closed = fclose(fp) == 0;
			fp = nullptr;
			if (!ok || !closed) throw FileAccessException("Could not write to file '" + filename_ + "'!");
		}
	}
	catch (...)
	{
		if (gz) gzclose(gz);
		if (fp) fclose(fp);
		std::lock_guard<std::mutex> g(mu_);
		if (!failure_) failure_ = std::current_exception();
	}
	cv_.notify_all();
}

void GzipTextWriter::close()
{
	if (closed_) return;
	{
		std::lock_guard<std::mutex> g(mu_);
		closing_ = true;
	}
	cv_.notify_all();
	if (writer_.joinable()) writer_.join();
	closed_ = true;
	// pieces still owned by pool tasks must not outlive this object: after a failure wait until they are all marked done
	{
		std::unique_lock<std::mutex> l(mu_);
		cv_.wait(l, [this] {
			for (const auto& p : queue_)
				if (!p->done) return false;
			return true;
		});
	}
	if (failure_) std::rethrow_exception(failure_);
}

} // namespace seqpurge
