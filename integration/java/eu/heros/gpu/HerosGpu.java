package eu.heros.gpu;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;
import static java.lang.foreign.ValueLayout.JAVA_SHORT;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout.PathElement;
import java.lang.foreign.MemorySegment;

import nl.tudelft.simulation.medlabs.common.MedlabsRuntimeException;

/**
 * One engine of libheros_gpu.so (include/heros_gpu.h) = one replication on one GPU. A handle is not thread-safe; the
 * reference drives its model from the single DSOL simulator thread, which is the thread that owns this object.
 * <p>
 * NOT COMPILED in the repository this file ships in (no JDK, medlabs:2.1.3 not vendored): it is the hand-written
 * layer a maintainer of medlabs-heros adds on top of the generated {@link HerosGpuNative}. The same layer, compiled and
 * tested against the CPU oracle, exists in C++ (include/heros_host.hpp) and Python (medlabs-heros_b200/disease.py).
 * </p>
 */
public final class HerosGpu implements AutoCloseable
{
    /** heros_model. */
    public static final int MODEL_AREA = 0, MODEL_DISTANCE = 1;

    /** heros_trigger: what the medlabs engine evaluates on (SURVEY.md contract row C8). */
    public static final int TRIGGER_EXIT_ONLY = 0, TRIGGER_ENTER_AND_EXIT = 1;

    /** heros_status. */
    public static final int OK = 0, ERR_INVALID = -1, ERR_CUDA = -2, ERR_UNSUPPORTED = -3, ERR_CAPACITY = -4,
            ERR_UNKNOWN_PARAMETER = -5;

    private static final int ABI_VERSION = 1;

    private static final long CFG_ABI = offset("abi_version"), CFG_MODEL = offset("model"), CFG_TRIGGER = offset("trigger"),
            CFG_DEVICE = offset("device"), CFG_SEED = offset("seed"), CFG_WINDOW = offset("window_hours"),
            CFG_FLAGS = offset("flags");

    /** Memory that lives as long as the engine. */
    final Arena arena = Arena.ofConfined();

    /** heros_handle. */
    final MemorySegment handle;

    private static long offset(final String field)
    {
        return HerosGpuNative.HEROS_CONFIG_LAYOUT.byteOffset(PathElement.groupElement(field));
    }

    /**
     * @param model int; MODEL_AREA or MODEL_DISTANCE (generic.diseasePropertiesModel, ConstructHerosModel.java:136-139)
     * @param trigger int; TRIGGER_EXIT_ONLY or TRIGGER_ENTER_AND_EXIT
     * @param device int; CUDA device ordinal
     * @param seed long; generic.Seed -- key of every draw
     * @param windowHours double; batch window, at most the latent period of the transmission model (24.0 is the usual
     *            choice: one call per simulated day)
     */
    public HerosGpu(final int model, final int trigger, final int device, final long seed, final double windowHours)
    {
        try
        {
            if ((int) HerosGpuNative.ABI_VERSION.invokeExact() != ABI_VERSION)
                throw new MedlabsRuntimeException("libheros_gpu.so has another ABI version than this binding");
            MemorySegment cfg = this.arena.allocate(HerosGpuNative.HEROS_CONFIG_LAYOUT);
            cfg.fill((byte) 0);
            cfg.set(JAVA_INT, CFG_ABI, ABI_VERSION);
            cfg.set(JAVA_INT, CFG_MODEL, model);
            cfg.set(JAVA_INT, CFG_TRIGGER, trigger);
            cfg.set(JAVA_INT, CFG_DEVICE, device);
            cfg.set(JAVA_LONG, CFG_SEED, seed);
            cfg.set(JAVA_DOUBLE, CFG_WINDOW, windowHours);
            cfg.set(JAVA_INT, CFG_FLAGS, 0);
            MemorySegment out = this.arena.allocate(ADDRESS);
            int status = (int) HerosGpuNative.CREATE.invokeExact(cfg, out);
            if (status != OK)
                throw new MedlabsRuntimeException("heros_create: status " + status + ": " + lastError(MemorySegment.NULL));
            this.handle = out.get(ADDRESS, 0);
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /** heros_last_error of this engine (of the last failed heros_create when h is NULL). */
    static String lastError(final MemorySegment h) throws Throwable
    {
        MemorySegment msg = (MemorySegment) HerosGpuNative.LAST_ERROR.invokeExact(h);
        return msg.equals(MemorySegment.NULL) ? "" : msg.reinterpret(4096).getString(0);
    }

    /** Turns a status other than HEROS_OK into the unchecked exception the reference's model code throws. */
    void check(final int status)
    {
        if (status == OK)
            return;
        String why;
        try
        {
            why = lastError(this.handle);
        }
        catch (Throwable t)
        {
            why = t.toString();
        }
        throw new MedlabsRuntimeException("libheros_gpu: status " + status + ": " + why);
    }

    /**
     * heros_set_location_types: the rows of locationtypes.csv in file order (ConstructHerosModel.java:230-269).
     * @param infectInSublocation boolean[]; column 3
     * @param correctionFactorArea double[]; column 5 (contagiousRateFactor)
     */
    public void setLocationTypes(final boolean[] infectInSublocation, final double[] correctionFactorArea)
    {
        try (Arena a = Arena.ofConfined())
        {
            int n = infectInSublocation.length;
            MemorySegment sub = a.allocate(n);
            for (int i = 0; i < n; i++)
                sub.set(JAVA_BYTE, i, (byte) (infectInSublocation[i] ? 1 : 0));
            MemorySegment corr = a.allocateFrom(JAVA_DOUBLE, correctionFactorArea);
            check((int) HerosGpuNative.SET_LOCATION_TYPES.invokeExact(this.handle, n, sub, corr, MemorySegment.NULL,
                    MemorySegment.NULL, MemorySegment.NULL));
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /**
     * heros_set_locations: location id = index (the ids of locations.csv are dense, ConstructHerosModel.java:300-395).
     * @param type byte[]; LocationType.getLocationTypeId()
     * @param nSub short[]; Location.getNumberOfSubLocations()
     * @param totalSurfaceM2 float[]; Location.getTotalSurfaceM2() = (float) (area * nb_sublocations), :381
     * @param lat float[]; latitude (only read by a device-side schedule; may be null)
     * @param lon float[]; longitude (may be null)
     */
    public void setLocations(final byte[] type, final short[] nSub, final float[] totalSurfaceM2, final float[] lat,
            final float[] lon)
    {
        try (Arena a = Arena.ofConfined())
        {
            check((int) HerosGpuNative.SET_LOCATIONS.invokeExact(this.handle, type.length, a.allocateFrom(JAVA_BYTE, type),
                    a.allocateFrom(JAVA_SHORT, nSub), a.allocateFrom(JAVA_FLOAT, totalSurfaceM2),
                    lat == null ? MemorySegment.NULL : a.allocateFrom(JAVA_FLOAT, lat),
                    lon == null ? MemorySegment.NULL : a.allocateFrom(JAVA_FLOAT, lon)));
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /**
     * heros_set_persons: person id = index (ConstructHerosModel.java:450-758).
     * @param age byte[]; Person.getAge()
     * @param female byte[]; 1 = female (Person.getGenderFemale())
     * @param role byte[]; social_role 1..15 of people.csv (selects the week pattern, HerosModel.java:484-504)
     * @param homeLocation int[]; home location id (device-side schedule only; may be null together with the next two)
     * @param homeSub short[]; home sub-location index
     * @param workLocation int[]; work / school location id, -1 = none
     */
    public void setPersons(final byte[] age, final byte[] female, final byte[] role, final int[] homeLocation,
            final short[] homeSub, final int[] workLocation)
    {
        try (Arena a = Arena.ofConfined())
        {
            boolean homes = homeLocation != null;
            check((int) HerosGpuNative.SET_PERSONS.invokeExact(this.handle, age.length, a.allocateFrom(JAVA_BYTE, age),
                    a.allocateFrom(JAVA_BYTE, female), a.allocateFrom(JAVA_BYTE, role),
                    homes ? a.allocateFrom(JAVA_INT, homeLocation) : MemorySegment.NULL,
                    homes ? a.allocateFrom(JAVA_SHORT, homeSub) : MemorySegment.NULL,
                    homes ? a.allocateFrom(JAVA_INT, workLocation) : MemorySegment.NULL));
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /**
     * The locations ConstructHerosModel.readLocationTable builds as LocationProbBased (the rows of
     * epidemiology/infection_rates.csv, ConstructHerosModel.java:274-295 and 382-388): location id, infection_rate_factor,
     * infection_rate (-1 = not given).  The engine infects in them by rate, never by occupancy (DESIGN.md 3.5).
     * @param location int[]; location ids as in locations.csv
     * @param rateFactor double[]; infectionRateFactor of every location
     * @param rate double[]; infectionRate of every location
     */
    public void setRateLocations(final int[] location, final double[] rateFactor, final double[] rate)
    {
        try (Arena a = Arena.ofConfined())
        {
            check((int) HerosGpuNative.SET_RATE_LOCATIONS.invokeExact(this.handle, location.length,
                    a.allocateFrom(JAVA_INT, location), a.allocateFrom(JAVA_DOUBLE, rateFactor),
                    a.allocateFrom(JAVA_DOUBLE, rate)));
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /**
     * heros_step: every window up to tUntil (bucket, transmit, draw, resolve, progress).
     * @param tUntil double; simulator time in hours
     */
    public void step(final double tUntil)
    {
        try
        {
            check((int) HerosGpuNative.STEP.invokeExact(this.handle, tUntil, MemorySegment.NULL));
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /**
     * heros_get_phase_counts: the DiseasePhase person counters, in the order the phases were registered
     * (Covid19Progression.java:130-137).
     * @return long[8]
     */
    public long[] phaseCounts()
    {
        try (Arena a = Arena.ofConfined())
        {
            MemorySegment c = a.allocate(JAVA_LONG, 8);
            check((int) HerosGpuNative.GET_PHASE_COUNTS.invokeExact(this.handle, c));
            return c.toArray(JAVA_LONG);
        }
        catch (RuntimeException e)
        {
            throw e;
        }
        catch (Throwable t)
        {
            throw new MedlabsRuntimeException(t);
        }
    }

    /** heros_destroy. */
    @Override
    public void close()
    {
        try
        {
            int ignored = (int) HerosGpuNative.DESTROY.invokeExact(this.handle);   // invokeExact: the int result must be taken
        }
        catch (Throwable t)
        {
            // nothing to do: the process is going down or the library has been unloaded
        }
        this.arena.close();
    }
}
