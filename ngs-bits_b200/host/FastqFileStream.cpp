#include "FastqFileStream.h"

#include <cstring>

namespace seqpurge
{

FastqFileStream::FastqFileStream(const std::string& filename) : filename_(filename)
{
	gz_ = gzopen(filename.c_str(), "rb");
	if (gz_ == nullptr) throw FileAccessException("Could not open file '" + filename + "' for reading!");
	gzbuffer(gz_, 128 * 1024); // zlib-internal buffer of the reference's FASTQ reader (128 x the 1 KiB line buffer)
}

FastqFileStream::~FastqFileStream()
{
	if (gz_) gzclose(gz_);
}

void FastqFileStream::readLine(std::string& out)
{
	out.clear();
	while (true)
	{
		char* s = gzgets(gz_, buffer_, (int)sizeof(buffer_));
		if (s == nullptr) // end of file, or an error such as a truncated gz stream
		{
			int error_no = Z_OK;
			const char* msg = gzerror(gz_, &error_no);
			if (error_no != Z_OK && error_no != Z_STREAM_END) throw FileParseException("Error while reading file '" + filename_ + "': " + msg);
			break;
		}
		out.append(s);
		if (!out.empty() && out.back() == '\n') break;
	}
	while (!out.empty() && (out.back() == '\n' || out.back() == '\r')) out.pop_back();
}

void FastqFileStream::readEntry(FastqEntry& entry)
{
	if (is_first_entry_)
	{
		readLine(last_output_);
		is_first_entry_ = false;
	}
	entry.header = last_output_;
	readLine(entry.bases);
	readLine(entry.header2);
	readLine(entry.qualities);
	readLine(last_output_);
}

FastqOutfileStream::FastqOutfileStream(const std::string& filename, int compression_level) : filename_(filename)
{
	gz_ = gzopen(filename.c_str(), "wb");
	if (gz_ == nullptr) throw FileAccessException("Could not open file '" + filename + "' for writing!");
	gzbuffer(gz_, 131072);
	if (compression_level < 0 || compression_level > 9)
		throw ArgumentException("Invalid gzip compression level '" + std::to_string(compression_level) + "' given for FASTQ file '" + filename + "'!");
	gzsetparams(gz_, compression_level, Z_DEFAULT_STRATEGY);
}

FastqOutfileStream::~FastqOutfileStream() { close(); }

void FastqOutfileStream::write(const FastqEntry& entry, size_t len)
{
	// the same byte stream as the reference's eight gzputs calls, assembled once per record
	line_.clear();
	line_.append(entry.header).push_back('\n');
	line_.append(entry.bases, 0, len).push_back('\n');
	line_.append(entry.header2).push_back('\n');
	line_.append(entry.qualities, 0, len).push_back('\n');
	if (gzwrite(gz_, line_.data(), (unsigned)line_.size()) != (int)line_.size()) throw FileAccessException("Could not write to file '" + filename_ + "'!");
}

void FastqOutfileStream::close()
{
	if (is_closed_) return;
	gzclose(gz_);
	is_closed_ = true;
}

} // namespace seqpurge
