// Standalone stand-in for the part of SuperTerrain+/SuperTerrain+/Public/SuperTerrain+/Exception/ that the single
// histogram filter can raise (STPFundamentalException.h:38-82, STPNumericDomainError.h:14-25, STPInvalidEnum.h:19-36,
// Utility/STPDeviceErrorHandler.hpp:22-25): same namespace, class names and inheritance, message-only constructors.
// Inside the reference tree the reference's own headers take this one's place.
#pragma once
#include <exception>
#include <string>

namespace SuperTerrainPlus::STPException {

	namespace STPFundamentalException {

		class STPBasic : public std::exception {
		private:

			std::string Message;

		public:

			explicit STPBasic(const std::string& description) : Message(description) { }

			~STPBasic() override = default;

			const char* what() const noexcept override { return this->Message.c_str(); }

		};

		class STPAssertion : public STPBasic {
		public:

			//"<expression>: <explanation>"
			STPAssertion(const char* expression, const std::string& explanation) :
				STPBasic(std::string(expression) + ": " + explanation) { }

		};

	}

	class STPNumericDomainError : public STPFundamentalException::STPAssertion {
	public:

		using STPAssertion::STPAssertion;

	};

	class STPInvalidEnum : public STPFundamentalException::STPBasic {
	public:

		using STPBasic::STPBasic;

	};

	class STPCUDAError : public STPFundamentalException::STPBasic {
	public:

		using STPBasic::STPBasic;

	};

	//additive: a shape the GPU kernels do not cover, or more than 2^32 bins in one chunk
	class STPUnsupportedOperation : public STPFundamentalException::STPBasic {
	public:

		using STPBasic::STPBasic;

	};

}
