#include "TextSource.h"

#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "GzipTextWriter.h"
#include "SeqPurgeTypes.h"

namespace seqpurge
{

namespace
{

class GzSource : public TextSource
{
public:
	explicit GzSource(const std::string& filename) : filename_(filename)
	{
		gz_ = gzopen(filename.c_str(), "rb");
		if (!gz_) throw FileAccessException("Could not open file '" + filename + "' for reading!");
		gzbuffer(gz_, 1 << 20);
	}
	GzSource(const std::string& filename, int fd) : filename_(filename) // takes the descriptor (positioned at a member start)
	{
		gz_ = gzdopen(fd, "rb");
		if (!gz_)
		{
			::close(fd);
			throw FileAccessException("Could not open file '" + filename + "' for reading!");
		}
		gzbuffer(gz_, 1 << 20);
	}
	~GzSource() override
	{
		if (gz_) gzclose(gz_);
	}
	size_t read(uint8_t* buf, size_t cap) override
	{
		const int n = gzread(gz_, buf, (unsigned)std::min<size_t>(cap, 1u << 30));
		if (n < 0)
		{
			int err = Z_OK;
			const char* msg = gzerror(gz_, &err);
			throw FileParseException("Error while reading file '" + filename_ + "': " + (msg ? msg : ""));
		}
		return (size_t)n;
	}

private:
	std::string filename_;
	gzFile gz_ = nullptr;
};

// size of the BGZF block that starts at p (n bytes available), 0 if p does not start one, -1 if more bytes are needed to tell
long bgzfBlockSize(const uint8_t* p, size_t n)
{
	if (n < 12) return -1;
	if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return 0;
	const size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8);
	if (n < 12 + xlen) return -1;
	size_t off = 12;
	while (off + 4 <= 12 + xlen)
	{
		const size_t slen = (size_t)p[off + 2] | ((size_t)p[off + 3] << 8);
		if (p[off] == 'B' && p[off + 1] == 'C' && slen == 2 && off + 6 <= 12 + xlen)
		{
			const size_t bs = ((size_t)p[off + 4] | ((size_t)p[off + 5] << 8)) + 1;
			// a block holds at least its header, the extra field and the 8-byte trailer (CRC32, ISIZE): anything smaller is corrupt and would
			// make the trailer reads of inflateJob point in front of the block
			return bs >= 12 + xlen + 8 ? (long)bs : 0;
		}
		off += 4 + slen;
	}
	return 0;
}

class BgzfSource : public TextSource
{
public:
	BgzfSource(const std::string& filename, WorkerPool* pool) : filename_(filename), pool_(pool)
	{
		fd_ = ::open(filename.c_str(), O_RDONLY);
		if (fd_ < 0) throw FileAccessException("Could not open file '" + filename + "' for reading!");
		feeder_ = std::thread([this]() { feed(); });
	}
	~BgzfSource() override
	{
		{
			std::lock_guard<std::mutex> g(mu_);
			abort_ = true;
		}
		cv_.notify_all();
		if (feeder_.joinable()) feeder_.join();
		// jobs still owned by pool tasks must not outlive this object
		std::unique_lock<std::mutex> l(mu_);
		cv_.wait(l, [this] {
			for (const auto& j : jobs_)
				if (!j->done) return false;
			return true;
		});
		if (fd_ >= 0) ::close(fd_);
	}
	size_t read(uint8_t* buf, size_t cap) override
	{
		for (;;)
		{
			if (tail_) return tail_->read(buf, cap);
			if (cur_ && cur_off_ < cur_->text.size())
			{
				const size_t n = std::min(cap, cur_->text.size() - cur_off_);
				memcpy(buf, cur_->text.data() + cur_off_, n);
				cur_off_ += n;
				return n;
			}
			cur_.reset();
			std::unique_lock<std::mutex> l(mu_);
			cv_.wait(l, [this] { return (!jobs_.empty() && jobs_.front()->done) || (jobs_.empty() && fed_) || failure_; });
			if (failure_) std::rethrow_exception(failure_);
			if (jobs_.empty())
			{
				if (fallback_offset_ >= 0) // the rest of the file is ordinary gzip
				{
					const int fd = ::dup(fd_);
					if (fd < 0 || ::lseek(fd, (off_t)fallback_offset_, SEEK_SET) < 0) throw FileAccessException("Could not read file '" + filename_ + "'!");
					fallback_offset_ = -1;
					l.unlock();
					tail_.reset(new GzSource(filename_, fd));
					continue;
				}
				return 0;
			}
			cur_ = std::move(jobs_.front());
			jobs_.pop_front();
			cur_off_ = 0;
			l.unlock();
			cv_.notify_all();
			if (cur_->error) throw FileParseException("Error while reading file '" + filename_ + "': corrupt BGZF block");
		}
	}

private:
	struct Job
	{
		std::vector<uint8_t> comp;                   // whole blocks
		std::vector<std::pair<uint32_t, uint32_t>> blocks; // offset in comp, size
		std::vector<uint8_t> text;
		bool done = false, error = false;
	};

	void inflateJob(Job* j)
	{
		bool error = false;
		try
		{
			size_t total = 0;
			for (const auto& b : j->blocks)
			{
				if (b.second < 12 + 8) throw Exception("corrupt BGZF block");
				const uint8_t* e = j->comp.data() + b.first + b.second;
				const size_t isize = (size_t)e[-4] | ((size_t)e[-3] << 8) | ((size_t)e[-2] << 16) | ((size_t)e[-1] << 24);
				if (isize > 65536) throw Exception("corrupt BGZF block"); // the format limits a block to 64 KiB of text
				total += isize;
			}
			j->text.resize(total);
			z_stream zs;
			memset(&zs, 0, sizeof(zs));
			if (inflateInit2(&zs, -15) != Z_OK) throw Exception("inflateInit2 failed");
			size_t out = 0;
			for (const auto& b : j->blocks)
			{
				const uint8_t* p = j->comp.data() + b.first;
				const uint8_t* e = p + b.second;
				const size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8);
				const size_t isize = (size_t)e[-4] | ((size_t)e[-3] << 8) | ((size_t)e[-2] << 16) | ((size_t)e[-1] << 24);
				const uint32_t want_crc = (uint32_t)e[-8] | ((uint32_t)e[-7] << 8) | ((uint32_t)e[-6] << 16) | ((uint32_t)e[-5] << 24);
				if (b.second < 12 + xlen + 8)
				{
					error = true;
					break;
				}
				if (isize == 0) continue; // e.g. the empty block that ends a BGZF file
				inflateReset(&zs);
				zs.next_in = const_cast<uint8_t*>(p + 12 + xlen);
				zs.avail_in = (uInt)(b.second - 12 - xlen - 8);
				zs.next_out = j->text.data() + out;
				zs.avail_out = (uInt)isize;
				const int rc = inflate(&zs, Z_FINISH);
				if (rc != Z_STREAM_END || zs.avail_out != 0 || (uint32_t)crc32(crc32(0L, Z_NULL, 0), j->text.data() + out, (uInt)isize) != want_crc)
				{
					error = true;
					break;
				}
				out += isize;
			}
			inflateEnd(&zs);
		}
		catch (...)
		{
			error = true;
		}
		std::vector<uint8_t>().swap(j->comp);
		{
			std::lock_guard<std::mutex> g(mu_);
			j->error = error;
			j->done = true;
			cv_.notify_all(); // under the lock: once `done` is visible the source may be destroyed, this task must not touch it afterwards
		}
	}

	void feed()
	{
		try
		{
			constexpr size_t kJobBytes = 512 << 10; // compressed bytes per inflate task
			const size_t max_jobs = 48;       // inflate tasks in flight (about 100 MB of text)
			std::vector<uint8_t> buf;
			size_t have = 0;    // valid bytes in buf
			long long base = 0; // file offset of buf[0]
			bool eof = false;
			for (;;)
			{
				// cut whole blocks off the front of buf into one job
				std::unique_ptr<Job> job(new Job());
				size_t off = 0;
				bool stop = false;
				while (off < have && off < kJobBytes)
				{
					const long bs = bgzfBlockSize(buf.data() + off, have - off);
					if (bs < 0 || (bs > 0 && off + (size_t)bs > have))
					{
						if (eof) throw FileParseException("Error while reading file '" + filename_ + "': truncated BGZF block");
						break; // need more bytes
					}
					if (bs == 0) // not a BGZF member: the rest goes through gzFile
					{
						stop = true;
						break;
					}
					job->blocks.emplace_back((uint32_t)off, (uint32_t)bs);
					off += (size_t)bs;
				}
				if (!job->blocks.empty())
				{
					job->comp.assign(buf.begin(), buf.begin() + (long)off);
					memmove(buf.data(), buf.data() + off, have - off);
					have -= off;
					base += (long long)off;
					Job* raw = job.get();
					{
						std::unique_lock<std::mutex> l(mu_);
						cv_.wait(l, [&] { return jobs_.size() < max_jobs || abort_; });
						if (abort_) break;
						jobs_.push_back(std::move(job));
					}
					pool_->run([this, raw]() { inflateJob(raw); });
				}
				if (stop)
				{
					std::lock_guard<std::mutex> g(mu_);
					fallback_offset_ = base;
					break;
				}
				if (eof && have == 0) break;
				if (!eof && have < 2 * kJobBytes)
				{
					buf.resize(std::max(buf.size(), have + 2 * kJobBytes));
					const ssize_t n = ::read(fd_, buf.data() + have, buf.size() - have);
					if (n < 0) throw FileAccessException("Could not read file '" + filename_ + "'!");
					if (n == 0) eof = true;
					have += (size_t)n;
				}
				{
					std::lock_guard<std::mutex> g(mu_);
					if (abort_) break;
				}
			}
		}
		catch (...)
		{
			std::lock_guard<std::mutex> g(mu_);
			if (!failure_) failure_ = std::current_exception();
		}
		{
			std::lock_guard<std::mutex> g(mu_);
			fed_ = true;
		}
		cv_.notify_all();
	}

	std::string filename_;
	WorkerPool* pool_;
	int fd_ = -1;
	std::thread feeder_;
	std::mutex mu_;
	std::condition_variable cv_;
	std::deque<std::unique_ptr<Job>> jobs_; // in file order
	bool fed_ = false, abort_ = false;
	long long fallback_offset_ = -1;
	std::exception_ptr failure_;
	std::unique_ptr<Job> cur_;
	size_t cur_off_ = 0;
	std::unique_ptr<TextSource> tail_;
};

} // namespace

bool isBgzf(const std::string& filename)
{
	FILE* f = fopen(filename.c_str(), "rb");
	if (!f) return false;
	uint8_t hdr[64];
	const size_t n = fread(hdr, 1, sizeof(hdr), f);
	fclose(f);
	return bgzfBlockSize(hdr, n) > 0;
}

std::unique_ptr<TextSource> openTextSource(const std::string& filename, WorkerPool* pool)
{
	if (pool && isBgzf(filename)) return std::unique_ptr<TextSource>(new BgzfSource(filename, pool));
	return std::unique_ptr<TextSource>(new GzSource(filename));
}

} // namespace seqpurge
