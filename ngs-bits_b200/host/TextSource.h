// TextSource.h -- inflated text of one input file for the stream pipeline.
//
// The reference reads FASTQ through gzFile (VersatileFile, src/cppCORE/VersatileFile.cpp:274-308): one thread inflates. That is what
// GzSource does (plain text, gzip, multi-member gzip). A BGZF file (the blocked gzip of htslib/bgzip, DRAGEN, samtools fastq: every
// member carries its compressed size in a 'BC' extra field and inflates on its own) is inflated by the shared worker pool instead:
// a feeder thread cuts the file into blocks, the pool inflates groups of blocks, read() hands the text out in file order
// (parallel inflate of blocked inputs). A member without the BC field in the middle of such a file makes
// the source fall back to gzFile for the rest of the file.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>

namespace seqpurge
{

class WorkerPool;

class TextSource
{
public:
	virtual ~TextSource() {}
	// up to cap bytes of text; 0 = end of file. Throws FileAccessException / FileParseException.
	virtual size_t read(uint8_t* buf, size_t cap) = 0;
};

// pool == nullptr (or a file that is not BGZF): serial gzFile reader
std::unique_ptr<TextSource> openTextSource(const std::string& filename, WorkerPool* pool);

// true if the file starts with a BGZF block
bool isBgzf(const std::string& filename);

} // namespace seqpurge
