// Output view of the single histogram filter; layout-identical to
// SuperTerrain+/SuperAlgorithm+/Host/Public/SuperAlgorithm+Host/STPSingleHistogram.hpp:15-50, so that
// SuperAlgorithm+Device's STPSingleHistogramWrapper::iterate reads it unchanged.
#pragma once
#include <SuperTerrain+/World/STPWorldMapPixelFormat.hpp>

namespace SuperTerrainPlus::STPAlgorithm {

	struct STPSingleHistogram {
	public:

		template<typename WT>
		struct STPGenericBin {
		public:

			//the sample value of this bin
			STPSample_t Item;
			//normalised weight (float) of the item in the pixel's window
			WT Weight;

		};
		typedef STPGenericBin<float> STPBin;

		//bins of all pixels, pixel after pixel in row-major order
		const STPBin* Bin;
		//index of each pixel's first bin; MapSize.x * MapSize.y + 1 entries, the last one is the number of bins
		const unsigned int* HistogramStartOffset;

	};

}
