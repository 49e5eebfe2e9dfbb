// FastqFileStream.h -- FASTQ(.gz) reader and writer with the observable behaviour of the reference's streams
// (src/cppNGS/FastqFileStream.{h,cpp}:119-193 over src/cppCORE/VersatileFile.cpp:286-308,395-414), Qt-free.
// Same zlib call sequence as the reference on both sides, so that output .gz files are byte-identical for identical records.
#pragma once
#include <zlib.h>

#include <string>

#include "SeqPurgeTypes.h"

namespace seqpurge
{

class FastqFileStream
{
public:
	explicit FastqFileStream(const std::string& filename);
	~FastqFileStream();
	FastqFileStream(const FastqFileStream&) = delete;
	FastqFileStream& operator=(const FastqFileStream&) = delete;

	bool atEnd() const { return gzeof(gz_) != 0; }
	void readEntry(FastqEntry& entry); // header / bases / header2 / qualities with one line of look-ahead
	const std::string& filename() const { return filename_; }

private:
	void readLine(std::string& out); // one line, trailing \n and \r removed
	std::string filename_;
	gzFile gz_;
	bool is_first_entry_ = true;
	std::string last_output_;
	char buffer_[1024];
};

class FastqOutfileStream
{
public:
	FastqOutfileStream(const std::string& filename, int compression_level);
	~FastqOutfileStream();
	FastqOutfileStream(const FastqOutfileStream&) = delete;
	FastqOutfileStream& operator=(const FastqOutfileStream&) = delete;

	// writes the first `len` bases/qualities of the entry (the trimmed record)
	void write(const FastqEntry& entry, size_t len);
	void close();

private:
	std::string filename_;
	gzFile gz_;
	bool is_closed_ = false;
	std::string line_;
};

} // namespace seqpurge
