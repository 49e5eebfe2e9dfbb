// fastq_dump -- test helper for the host-side FASTQ streams: prints what FastqFileStream::readEntry delivers, one line per call
// ("<atEnd before the call>\t<header>\t<bases>\t<header2>\t<qualities>"), and optionally copies the entries through
// FastqOutfileStream. Used by tests/test_fastq_streams.py against the reference's reader/writer fixtures
// (src/cppNGS-TEST/FastqFileStream_Test.cpp:130-452).
#include <iostream>
#include <memory>

#include "FastqFileStream.h"

using namespace seqpurge;

int main(int argc, char** argv)
{
	if (argc < 2)
	{
		std::cerr << "usage: fastq_dump <in.fastq[.gz]> [out.fastq.gz]" << std::endl;
		return 2;
	}
	try
	{
		FastqFileStream in(argv[1]);
		std::unique_ptr<FastqOutfileStream> out;
		if (argc > 2) out.reset(new FastqOutfileStream(argv[2], 1));
		FastqEntry e;
		int calls = 0;
		while (true)
		{
			const bool at_end = in.atEnd();
			in.readEntry(e);
			std::cout << (at_end ? 1 : 0) << '\t' << e.header << '\t' << e.bases << '\t' << e.header2 << '\t' << e.qualities << '\n';
			if (at_end || ++calls > 100000) break;
			if (out) out->write(e, e.bases.size());
		}
		std::cout << "END\t" << (in.atEnd() ? 1 : 0) << std::endl;
		return 0;
	}
	catch (const FileParseException& ex)
	{
		std::cout << "FileParseException\t" << ex.what() << std::endl;
		return 3;
	}
	catch (const std::exception& ex)
	{
		std::cout << "Exception\t" << ex.what() << std::endl;
		return 1;
	}
}
